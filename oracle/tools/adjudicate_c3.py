"""Adjudication of BASELINE configs[2] on the Bethe cluster graph (TEST INFRASTRUCTURE).

The C3 workload (muller_2022, Bethe cluster graph, regularizebeliefs_bycluster!, 10 loopy iterations over two
spanning trees, src/clustergraphbeliefs.jl:235-249, src/beliefupdates.jl:55-83) is ill-conditioned: the 800
factor-less variable clusters get eps = eps(Float64).  This script measures how far every double-precision
restatement is from the EXACT answer of the reference's algorithm on the same binary64 inputs:

  exact   = oracle/c compiled in IEEE binary128 (113 bits; oracle/c/build.py quad=True), cross-checked on one
            replicate by an independent mpmath (200-bit) run of the same message sequence (--mpmath)
  twin    = oracle/c in binary64 (LAPACK-order Cholesky, X_invA_Xt formulation of the reference)
  product = the product's kernel bodies (libpgbp_emul.so = the CUDA kernel bodies compiled as host C++ with
            contraction off and explicit fma(): bit-identical arithmetic to the GPU up to libm's log/sqrt)

and an empirical condition estimate: the change of the exact answer under a relative perturbation of 2^-53 of
the inputs (tip data), i.e. what ANY backward-stable double-precision algorithm may legitimately differ by.

    python oracle/tools/adjudicate_c3.py [--n 4] [--niter 10] [--workload c3|c3l] [--mpmath]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_product(w, params, tips, niter, lib):
    import pgbp_b200
    d = w.d
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, tips.shape[0])
    bt.assignfactors(params, tips)
    bt.regularizebeliefs_bycluster()
    succ, iscal = bt.calibrate(None, niter)
    assert succ.all()
    return bt.factored_energy()


def run_mpmath(w, params, tip, niter, dps=60):
    """Independent high-precision run (mpmath, `dps` digits) of assignfactors! (univariate BM, fixed root) +
    regularizebeliefs_bycluster! + calibrate!(niter) + factored_energy on ONE replicate, straight from the
    plan arrays; written against the reference (src/beliefs.jl:786-861, src/clustergraphbeliefs.jl:235-275,
    src/beliefupdates.jl:55-83,634-665, src/score.jl:162-182), sharing no code with oracle/c."""
    import mpmath as mp
    mp.mp.dps = dps
    d = w.d
    assert d["ntraits"] == 1 and d["families"]["root_fixed"] == 1
    nc, dims = d["nclusters"], d["belief_dim"]
    nb = len(dims)
    sigma2, mu = mp.mpf(float(params[0, 0])), mp.mpf(float(params[0, 1]))
    J = [mp.zeros(m, m) if m else mp.matrix(0, 0) for m in dims]
    h = [mp.zeros(m, 1) if m else mp.matrix(0, 1) for m in dims]
    g = [mp.mpf(0) for _ in dims]
    F = d["families"]
    log2pi = mp.log(2 * mp.pi)
    for v in range(F["nnodes"]):
        c = F["node_cluster"][v]
        k0, k1 = F["mem_off"][v], F["mem_off"][v + 1]
        nm = k1 - k0
        if nm == 1:
            continue  # fixed root: no prior factor
        t0 = mp.mpf(0)
        if nm == 2:
            t0 = mp.mpf(F["mem_length"][k0 + 1])
        else:
            for k in range(k0 + 1, k1):
                t0 += mp.mpf(F["mem_gamma"][k]) ** 2 * mp.mpf(F["mem_length"][k])
        j = 1 / (sigma2 * t0)
        gg = -(log2pi + mp.log(sigma2 * t0)) / 2
        cf = [mp.mpf(1)] + [(-mp.mpf(1) if nm == 2 else -mp.mpf(F["mem_gamma"][k0 + a])) for a in range(1, nm)]
        JJ = [[cf[a] * cf[b] * j for b in range(nm)] for a in range(nm)]
        hh = [mp.mpf(0)] * nm
        gone = [False] * nm
        for a in range(nm):
            if F["mem_pos"][k0 + a] >= 0:
                continue
            y = mp.mpf(float(tip[F["node_datarow"][v], 0])) if a == 0 else mu
            gg += hh[a] * y - y * JJ[a][a] * y / 2
            for r in range(nm):
                if not gone[r] and r != a:
                    hh[r] -= JJ[r][a] * y
            gone[a] = True
        for a in range(nm):
            pa = F["mem_pos"][k0 + a]
            if pa < 0:
                continue
            h[c][pa] += hh[a]
            for b in range(nm):
                pb = F["mem_pos"][k0 + b]
                if pb >= 0:
                    J[c][pa, pb] += JJ[a][b]
        g[c] += gg
    fJ = [x.copy() for x in J[:nc]]
    fh = [x.copy() for x in h[:nc]]
    fg = list(g[:nc])
    sep = d["sepset_clusters"]
    up = d["upind"]
    # regularizebeliefs_bycluster!
    nbrs = [[] for _ in range(nc)]
    for jx, (a, b) in enumerate(sep):
        nbrs[a].append((b, jx)); nbrs[b].append((a, jx))
    eps0 = mp.mpf(2.220446049250313e-16)
    for c in range(nc):
        m = dims[c]
        eps = eps0
        for r in range(m):
            for q in range(m):
                eps = max(eps, abs(J[c][r, q]))
        for other, jx in sorted(nbrs[c]):
            s = dims[nc + jx]
            side = 0 if sep[jx][0] == c else 1
            for k in range(s):
                u = up[jx][side][k]
                J[c][u, u] += eps
                J[nc + jx][k, k] += eps
    sepof = {frozenset(ab): jx for jx, ab in enumerate(sep)}

    def propagate(frm, jx, to):
        side_f = 0 if sep[jx][0] == frm else 1
        keep = list(up[jx][side_f]); upT = list(up[jx][1 - side_f])
        m = dims[frm]; s = dims[nc + jx]
        integ = [v for v in range(m) if v not in keep]
        ni = len(integ)
        mJ = mp.matrix(s, s); mh = mp.matrix(s, 1)
        for r in range(s):
            mh[r] = h[frm][keep[r]]
            for q in range(s):
                mJ[r, q] = J[frm][keep[r], keep[q]]
        mg = g[frm]
        if ni:
            Ji = mp.matrix(ni, ni); hi = mp.matrix(ni, 1); Jki = mp.matrix(s, ni)
            for r in range(ni):
                hi[r] = h[frm][integ[r]]
                for q in range(ni):
                    Ji[r, q] = J[frm][integ[r], integ[q]]
                for q in range(s):
                    Jki[q, r] = J[frm][keep[q], integ[r]]
            allzero = all(abs(x) <= eps0 for x in list(Ji) + list(hi) + list(Jki))
            if not allzero:
                Jinv = Ji ** -1
                mui = Jinv * hi
                mJ = mJ - Jki * Jinv * Jki.T
                mh = mh - Jki * mui
                mg = mg + (ni * log2pi - mp.log(mp.det(Ji)) + (hi.T * mui)[0]) / 2
        sb = nc + jx
        for r in range(s):
            for q in range(s):
                dJ = mJ[r, q] - J[sb][r, q]
                J[sb][r, q] = mJ[r, q]
                J[to][upT[r], upT[q]] += dJ
            dh = mh[r] - h[sb][r]
            h[sb][r] = mh[r]
            h[to][upT[r]] += dh
        dg = mg - g[sb]
        g[sb] = mg
        g[to] += dg
    for it in range(niter):
        for par, chi in d["trees"]:
            n = len(par)
            for i in range(n - 1, -1, -1):
                propagate(chi[i], sepof[frozenset((par[i], chi[i]))], par[i])
            for i in range(n):
                propagate(par[i], sepof[frozenset((par[i], chi[i]))], chi[i])
    en = mp.mpf(0); ent = mp.mpf(0)
    for c in range(nc):
        m = dims[c]
        if m == 0:
            en -= fg[c]; continue
        S = J[c] ** -1
        mu_c = S * h[c]
        tr = sum((fJ[c] * S)[k, k] for k in range(m))
        en += (tr + (mu_c.T * fJ[c] * mu_c)[0]) / 2 - (fh[c].T * mu_c)[0] - fg[c]
        ent += (m * (log2pi + 1) - mp.log(mp.det(J[c]))) / 2
    for jx in range(len(sep)):
        m = dims[nc + jx]
        if m:
            ent -= (m * (log2pi + 1) - mp.log(mp.det(J[nc + jx]))) / 2
    return np.array([float(en), float(ent), float(-(en - ent))])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--niter", type=int, default=10)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--mpmath", action="store_true")
    ap.add_argument("--lib", default=None, help="product library (default: the host-emulation build of the kernel bodies)")
    args = ap.parse_args()
    import bench
    from harness import get_lib
    import pgbp_b200
    from oracle.cport import COracle
    w = bench.WORKLOADS[args.workload]()
    params, tips = w.inputs(args.n, 0)
    co = COracle.from_plan_dict(w.d)
    kw = dict(w.cpu_kw, niter=args.niter, root_belief=w.d["root_cluster"], want_fe=True)
    exact = co.run_batch(params, tips, quad=True, **kw)["fe"]
    twin = co.run_batch(params, tips, **kw)["fe"]
    lib = pgbp_b200.Library(args.lib) if args.lib else get_lib("emul")
    prod = run_product(w, params, tips, args.niter, lib)
    # condition estimate: exact answer under input perturbations of relative size 2^-53 (unit roundoff)
    rng = np.random.default_rng(1)
    cond = np.zeros(3)
    for _ in range(4):
        tp = tips * (1 + rng.choice([-1.0, 1.0], size=tips.shape) * 2.0 ** -53)
        pert = co.run_batch(params, tp, quad=True, **kw)["fe"]
        cond = np.maximum(cond, np.max(np.abs(pert / exact - 1), axis=0))
    rel = lambda a: np.max(np.abs(a / exact - 1), axis=0)  # noqa: E731
    out = {"workload": args.workload, "replicates": args.n, "niter": args.niter,
           "columns": ["average energy", "approximate entropy", "factored energy"],
           "twin_vs_exact": rel(twin).tolist(), "product_vs_exact": rel(prod).tolist(),
           "product_vs_twin": np.max(np.abs(prod / twin - 1), axis=0).tolist(),
           "exact_sensitivity_to_1ulp_input_perturbation": cond.tolist(),
           "condition_estimate": (cond / 2.0 ** -53).tolist()}
    if args.mpmath:
        mpv = run_mpmath(w, params, tips[0], args.niter)
        out["mpmath_vs_quad_replicate0"] = np.abs(mpv / exact[0] - 1).tolist()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
