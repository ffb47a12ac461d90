"""Adjudication of the ill-conditioned loopy configuration (BASELINE configs[2] on the Bethe cluster graph) and the
reference-order validation mode.

Which side is closer to the exact answer of the reference's algorithm?  "Exact" = the C oracle compiled in IEEE
binary128 (oracle/c, -DPGBPO_QUAD, 113 bits), itself cross-checked here by an independent mpmath run (60 digits)
written straight from the reference (oracle/tools/adjudicate_c3.py).  Findings asserted below (also in DESIGN.md
section 2 and profiles/r2_c3_bethe_adjudication.json):
  * the exact result is insensitive to 1-ulp perturbations of the inputs (condition estimate ~ 1): the PROBLEM is
    well conditioned, the reference's ALGORITHM evaluated in binary64 is not -- its own formulation (oracle/c in
    binary64) ends ~3e-5 away from exact after 10 iterations, the product's fused formulation ~4.5e-5;
  * with PGBP_CAL_REFORDER (every message in the reference's LAPACK-style operation order) the product's J and h of
    EVERY belief are bit-identical to the binary64 twin: the difference is rounding order, nothing else.
Tolerance stated for configs[2]-Bethe: 1e-4 relative against the exact (binary128) factored energy.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tools"))

import bench  # noqa: E402
import pgbp_b200  # noqa: E402
from harness import BACKENDS, get_lib  # noqa: E402
from oracle.cport import COracle  # noqa: E402

import adjudicate_c3 as ADJ  # noqa: E402

BETHE_TOL = 1e-4  # stated tolerance of configs[2]-Bethe (relative, against the exact factored energy)


def _plan(w, lib):
    d = w.d
    return pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)


def test_quad_build_agrees_with_double_on_well_posed_configs():
    # exact clique-tree calibration (C2) and the LTRIP loopy graph: binary64 and binary128 runs of the same code agree
    # to the clique-tree tolerance -- the quad build is the same algorithm, only wider
    for key, tol in (("c2", 1e-12), ("c3l", 1e-10)):
        w = bench.WORKLOADS[key]()
        params, tips = w.inputs(8, 0)
        co = COracle.from_plan_dict(w.d)
        kw = dict(w.cpu_kw, root_belief=w.d["root_cluster"], want_fe=True)
        a = co.run_batch(params, tips, **kw)
        q = co.run_batch(params, tips, quad=True, **kw)
        assert np.array_equal(a["iscal"], q["iscal"])
        assert np.max(np.abs(a["fe"] / q["fe"] - 1)) <= tol, key
        assert np.max(np.abs(a["loglik"] / q["loglik"] - 1)) <= tol, key


def test_mpmath_confirms_the_quad_oracle_on_the_bethe_graph():
    # independent 60-digit run of assignfactors! + regularizebeliefs_bycluster! + 2 loopy iterations + free energy
    w = bench.C3()
    params, tips = w.inputs(1, 0)
    kw = dict(w.cpu_kw, niter=2, root_belief=w.d["root_cluster"], want_fe=True)
    quad = COracle.from_plan_dict(w.d).run_batch(params, tips, quad=True, **kw)["fe"][0]
    mpv = ADJ.run_mpmath(w, params, tips[0], 2)
    assert np.max(np.abs(mpv / quad - 1)) <= 1e-15


def test_bethe_adjudication_twin_and_product_vs_exact():
    w = bench.C3()
    n = 6
    params, tips = w.inputs(n, 0)
    co = COracle.from_plan_dict(w.d)
    kw = dict(w.cpu_kw, root_belief=w.d["root_cluster"], want_fe=True)  # niter = 10
    exact = co.run_batch(params, tips, quad=True, **kw)["fe"]
    twin = co.run_batch(params, tips, **kw)["fe"]
    prod = ADJ.run_product(w, params, tips, w.niter, get_lib("emul"))
    e_twin = np.max(np.abs(twin[:, 2] / exact[:, 2] - 1))
    e_prod = np.max(np.abs(prod[:, 2] / exact[:, 2] - 1))
    # both double-precision evaluations sit 1e-5 .. 1e-4 away from exact: far above 1e-10, inside the stated tolerance
    assert 1e-7 < e_twin <= BETHE_TOL and 1e-7 < e_prod <= BETHE_TOL
    assert e_prod <= 3 * e_twin  # the fused formulation is not worse than the reference's by more than a small factor
    # the problem itself is well conditioned: 1-ulp input perturbations move the exact answer by a few ulp
    rng = np.random.default_rng(1)
    tp = tips * (1 + rng.choice([-1.0, 1.0], size=tips.shape) * 2.0 ** -53)
    pert = co.run_batch(params, tp, quad=True, **kw)["fe"]
    assert np.max(np.abs(pert / exact - 1)) <= 1e-14


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("key", ["c3", "c3l"])
def test_reference_order_mode_is_bit_identical_to_the_twin(backend, key):
    # PGBP_CAL_REFORDER: J and h of every belief after 10 loopy iterations equal the binary64 twin bit for bit
    # (g to 1e-13: the GPU's log differs from glibc's in the last place), calibration flags equal
    lib = get_lib(backend)
    w = bench.WORKLOADS[key]()
    d = w.d
    n = 8
    params, tips = w.inputs(n, 0)
    co = COracle.from_plan_dict(d)
    twin = co.run_batch(params, tips, root_belief=d["root_cluster"], want_fe=True, want_state=True, **w.cpu_kw)
    bt = pgbp_b200.BatchedClusterGraphBelief(_plan(w, lib), n)
    bt.assignfactors(params, tips)
    bt.regularizebeliefs_bycluster()
    succ, iscal = bt.calibrate(None, w.niter, reference_order=True)
    assert succ.all() and np.array_equal(iscal, twin["iscal"])
    rng = np.random.default_rng(4)
    nb = len(d["belief_dim"])
    for j in rng.choice(nb, size=160, replace=False):
        m = d["belief_dim"][j]
        if m == 0:
            continue
        J, h, g = bt.get_belief(int(j) + 1)
        iu = np.triu_indices(m)
        for e in range(n):
            Jt, ht, gt = co.unpack(twin["state"][e], int(j))
            assert np.array_equal(J[e][iu], Jt[iu]) and np.array_equal(h[e], ht), (j, e)
            assert abs(g[e] - gt) <= 1e-12 * max(1.0, abs(gt))
    fe = bt.factored_energy()
    assert np.max(np.abs(fe / twin["fe"] - 1)) <= 1e-10  # same beliefs, two orders of the energy sums
    # and the default (fused) formulation on the same inputs: inside the stated tolerance of the configuration
    bt.init_beliefs_reset_fromfactors()
    bt.init_messagecalibrationflags_reset()
    bt.regularizebeliefs_bycluster()
    succ, iscal = bt.calibrate(None, w.niter)
    assert succ.all() and np.array_equal(iscal, twin["iscal"])
    fe2 = bt.factored_energy()
    assert np.max(np.abs(fe2[:, 2] / twin["fe"][:, 2] - 1)) <= (BETHE_TOL if key == "c3" else 1e-10)
