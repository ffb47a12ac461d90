"""Host-side drivers above the message-passing path, batched over data sets.

Mirrors of the reference's L4 drivers (src/calibration.jl:163-517) that only orchestrate calls of the
C ABI; they contain no belief arithmetic of their own beyond the closed-form REML sums the reference
also does on the host.
"""
from __future__ import annotations

import numpy as np

from .api import BatchedClusterGraphBelief, ClusterGraphPlan, bm_params


def calibrate_exact_cliquetree(plan_improper: ClusterGraphPlan, plan_fixed: ClusterGraphPlan, spt_improper, spt_fixed,
                               tipdata, device: int = 0):
    """calibrate_exact_cliquetree! (src/calibration.jl:404-517) for B data sets at once.

    Exact REML estimate of the Brownian-motion rate matrix and ML estimate of the root mean on a clique
    tree: (1) calibrate with identity rate and improper (infinite-variance) root prior; (2) mu_hat = root
    mean; sigma2_hat = sum_v E[(x_v - sum_k gamma_k x_pa_k)(..)'] / t_v  over  sum_v (1 - Var(..)/t_v)
    from the conditional moments of every node family (integratebelief! + inv(J) of its cluster);
    (3) re-assign factors at (sigma2_hat, mu_hat, fixed root), calibrate, integrate -> log-likelihood.

    plan_improper / plan_fixed: plans of the same clique tree allocated for a random root (root in scope)
    and for a fixed root, both with a node-family table; spt_*: their spanning tree (reference 4-tuple or
    tree id).  tipdata: (B, ntips, p), no missing values.
    Returns (sigma2_hat [B,p,p], mu_hat [B,p], loglik [B])."""
    td = np.ascontiguousarray(np.asarray(tipdata, dtype=float))
    if td.ndim == 2:
        td = td[None]
    B, ntips, p = td.shape
    fam = plan_improper.families
    if fam is None or plan_fixed.families is None:
        raise ValueError("both plans need a node-family table")
    # ---- (1) calibration parameters: R = I, mu = 0, v = Inf (src/calibration.jl:424)
    bb = BatchedClusterGraphBelief(plan_improper, B, device=device, factors=False, residuals=True)
    bb.assignfactors(bm_params([np.eye(p)], np.zeros(p), np.diag([np.inf] * p)), td)
    succ, _ = bb.calibrate([spt_improper] if not isinstance(spt_improper, (int, np.integer)) else [int(spt_improper)])
    ok = succ.copy()
    # ---- (2) conditional moments per node family (:447-499)
    off, pos = fam["mem_off"], fam["mem_pos"]
    length, gamma, ncl, row = fam["mem_length"], fam["mem_gamma"], fam["node_cluster"], fam["node_datarow"]
    moments = {}

    def mom(c):
        if c not in moments:
            mu, cov, _ = bb.integratebelief_cov(c + 1)
            moments[c] = (mu, cov)
        return moments[c]
    rootpos = pos[off[0]]
    if rootpos < 0:
        raise ValueError("plan_improper must have the root in scope")
    mu_hat = mom(ncl[0])[0][:, rootpos:rootpos + p].copy()
    num = np.zeros((B, p, p))
    den = np.zeros(B)
    for v in range(1, fam["nnodes"]):
        k0, k1 = off[v], off[v + 1]
        par = range(k0 + 1, k1)
        t = sum(gamma[k] ** 2 * length[k] for k in par)
        if t == 0.0:
            continue  # 0 length => variance parameter absent from the factor
        mu, cov = mom(ncl[v])
        if row[v] >= 0:  # tip: one parent, data clamped
            pa = pos[k0 + 1]
            if pa < 0:
                continue
            d = mu[:, pa:pa + p] - td[:, row[v], :]
            num += d[:, :, None] * d[:, None, :] / t
            den += 1 - cov[:, pa, pa] / t  # assumes inv(J) proportional to the rate matrix
        else:
            ch = pos[k0]
            d = mu[:, ch:ch + p].copy()
            dv = cov[:, ch, ch].copy()
            for k in par:
                d -= gamma[k] * mu[:, pos[k]:pos[k] + p]
                dv -= 2 * gamma[k] * cov[:, ch, pos[k]]
                for k2 in par:
                    dv += gamma[k] * gamma[k2] * cov[:, pos[k], pos[k2]]
            num += d[:, :, None] * d[:, None, :] / t
            den += 1 - dv / t
    sigma2_hat = num / den[:, None, None]
    del bb
    # ---- (3) likelihood at the optimum: fixed root at mu_hat (:501-516)
    bf = BatchedClusterGraphBelief(plan_fixed, B, device=device, factors=False, residuals=True)
    params = np.stack([bm_params([sigma2_hat[e]], mu_hat[e]) for e in range(B)])
    bf.assignfactors(params, td)
    sf = [spt_fixed] if not isinstance(spt_fixed, (int, np.integer)) else [int(spt_fixed)]
    succ, _ = bf.calibrate(sf)
    ok &= succ
    rootj = plan_fixed.trees[plan_fixed.tree_id(sf[0])][0][0] + 1
    loglik = bf.integratebelief(rootj, want_mu=False)[1]
    loglik[~ok] = np.nan
    return sigma2_hat, mu_hat, loglik


def _bm_transforms(model, p, v):
    """params_optimize / params_original of the Brownian-motion models
    (src/evomodels/homogeneousbrownianmotion.jl:48-49, 89-90, 130-159): log-rates (log-Cholesky factor for the full
    multivariate model) and root means."""
    V = np.zeros((p, p)) if v is None else np.atleast_2d(np.asarray(v, dtype=float))
    if model == "UnivariateBrownianMotion":
        if p != 1:
            raise ValueError("UnivariateBrownianMotion needs 1 trait")
        return (lambda s2, mu: np.array([np.log(s2), float(np.ravel(mu)[0])]),
                lambda th: bm_params([np.exp(th[0])], [th[1]], V))
    if model == "MvDiagBrownianMotion":
        return (lambda R, mu: np.concatenate([np.log(np.asarray(R, float)), np.asarray(mu, float)]),
                lambda th: bm_params([np.exp(th[:p])], th[p:], V))
    if model == "MvFullBrownianMotion":
        # log-Cholesky parametrisation (src/evomodels/homogeneousbrownianmotion.jl:130-159): R = U'U, the log of
        # U's diagonal first, then the entries above the diagonal column by column, then the root mean
        iu = [(i, j) for j in range(1, p) for i in range(j)]

        def to_opt(R, mu):
            U = np.linalg.cholesky(np.asarray(R, float)).T
            return np.concatenate([np.log(np.diag(U)), [U[i, j] for i, j in iu], np.asarray(mu, float)])

        def to_orig(th):
            U = np.zeros((p, p))
            U[np.arange(p), np.arange(p)] = np.exp(th[:p])
            for k, (i, j) in enumerate(iu):
                U[i, j] = th[p + k]
            return bm_params([U.T @ U], th[p + len(iu):], V)
        return to_opt, to_orig
    raise ValueError("model must be UnivariateBrownianMotion, MvDiagBrownianMotion or MvFullBrownianMotion")


def calibrate_optimize_cliquetree(plan: ClusterGraphPlan, spt, tipdata, model="UnivariateBrownianMotion", start=(1.0, 0.0),
                                  v=None, maxiter=30, fd_step=1e-6, device: int = 0):
    """calibrate_optimize_cliquetree! (src/calibration.jl:182-234): maximum likelihood of the BM rate(s) and
    root mean(s) on a clique tree, objective = assignfactors! + postorder + integratebelief!(root).

    The reference evaluates its objective (and, through Optim's finite differences, its gradient) one
    parameter vector at a time; here every L-BFGS iteration evaluates the whole central-difference
    stencil (2n+1 parameter vectors) in ONE batched device call.  Failed Choleskys give +Inf, as in
    the reference (:196-219).  Returns (theta_hat in original parametrisation as a bm_params record,
    loglik, scipy result)."""
    from scipy.optimize import minimize
    td = np.ascontiguousarray(np.asarray(tipdata, dtype=float))
    if td.ndim == 2:
        td = td[None]
    p = td.shape[2]
    to_opt, to_orig = _bm_transforms(model, p, v)
    theta0 = to_opt(*start)
    n = theta0.size
    bb = BatchedClusterGraphBelief(plan, 2 * n + 1, device=device, factors=False, residuals=False)
    tid = plan.tree_id(spt)
    rootj = plan.trees[tid][0][0] + 1

    def scores(thetas):
        bb.clear_status()
        bb.assignfactors(np.stack([to_orig(t) for t in thetas]), td)
        succ = bb.propagate_1traversal_postorder(tid, update_residualnorm=False)
        ll = bb.integratebelief(rootj, want_mu=False)[1]
        out = -ll
        out[~succ | ~np.isfinite(ll) | (bb.status() != 0)] = np.inf
        return out

    def fun(theta):
        st = np.tile(theta, (2 * n + 1, 1))
        for k in range(n):
            st[1 + 2 * k, k] += fd_step
            st[2 + 2 * k, k] -= fd_step
        f = scores(st)
        return f[0], (f[1::2] - f[2::2]) / (2 * fd_step)

    res = minimize(fun, theta0, jac=True, method="L-BFGS-B", options=dict(maxiter=maxiter, ftol=1e-14, gtol=1e-9))
    return to_orig(res.x), -float(res.fun), res


def calibrate_optimize_clustergraph(plan: ClusterGraphPlan, schedule, tipdata, model="UnivariateBrownianMotion",
                                    start=(1.0, 0.0), v=None, maxiter=100, regfun="bycluster", optim_iterations=30,
                                    fd_step=1e-6, device: int = 0):
    """calibrate_optimize_clustergraph! (src/calibration.jl:309-359): maximise the factored energy (the negative
    Bethe free energy; the log-likelihood when the cluster graph is a clique tree) over the BM rate(s) and
    root mean(s) on an arbitrary cluster graph.

    Objective, as in the reference (:323-351): assignfactors! -> factor snapshot -> reset of the message
    residual flags -> regularisation (`regfun`: "bycluster" | "onschedule" | a node-subtree program for
    regularizebeliefs_bynodesubtree!) -> calibrate!(schedule, maxiter, auto=true) -> free_energy[3];
    +Inf when a message fails.  `schedule` = spanningtrees_clusterlist of the graph (reference 4-tuples or
    tree ids; None = every tree of the plan).  Every L-BFGS iteration evaluates its central-difference
    stencil (2n+1 parameter vectors) as ONE batch; auto-stop is per batch element.
    Returns (theta_hat as a bm_params record, factored energy, scipy result)."""
    from scipy.optimize import minimize
    td = np.ascontiguousarray(np.asarray(tipdata, dtype=float))
    if td.ndim == 2:
        td = td[None]
    p = td.shape[2]
    to_opt, to_orig = _bm_transforms(model, p, v)
    theta0 = to_opt(*start)
    n = theta0.size
    bb = BatchedClusterGraphBelief(plan, 2 * n + 1, device=device, factors=True, residuals=True)

    def scores(thetas):
        bb.clear_status()
        bb.assignfactors(np.stack([to_orig(t) for t in thetas]), td)  # also snapshots the factors
        bb.init_messagecalibrationflags_reset(True)
        if regfun == "bycluster":
            bb.regularizebeliefs_bycluster()
        elif regfun == "onschedule":
            bb.regularizebeliefs_onschedule()
        elif regfun is not None:
            bb.regularizebeliefs_bynodesubtree(regfun)
        succ, _ = bb.calibrate(schedule, maxiter, auto=True)
        fe = bb.free_energy()[:, 2]
        out = fe.copy()
        out[~succ | ~np.isfinite(fe) | (bb.status() != 0)] = np.inf
        return out

    def fun(theta):
        st = np.tile(theta, (2 * n + 1, 1))
        for k in range(n):
            st[1 + 2 * k, k] += fd_step
            st[2 + 2 * k, k] -= fd_step
        f = scores(st)
        return f[0], (f[1::2] - f[2::2]) / (2 * fd_step)

    res = minimize(fun, theta0, jac=True, method="L-BFGS-B", options=dict(maxiter=optim_iterations, ftol=1e-14, gtol=1e-9))
    return to_orig(res.x), -float(res.fun), res
