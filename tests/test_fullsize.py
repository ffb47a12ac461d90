"""GPU parity at BASELINE.json's full sizes, through the C ABI.

The NumPy oracle cannot run 65,536 replicates or a 21,999-node network in seconds, so at these
sizes the checks are (i) the C/OpenMP twin of the oracle (oracle/c, itself pinned against the NumPy
oracle in test_parity / test_synth) on every element or a large sample, and (ii) size-independent
properties of a calibrated clique tree: every belief integrates to the same log-likelihood
(test/test_calibration.jl:55-58), factored energy == log-likelihood, a second calibration is a
fixed point (all residuals below tolerance), and a repeated run is bit-identical."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import pgbp_b200  # noqa: E402
from harness import get_lib, relerr  # noqa: E402
from oracle.cport import COracle  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-10


def plan_of(w, lib):
    d = w.d
    return pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)


def test_c2_full_batch_properties_and_cport():
    lib = get_lib("cuda")
    w = bench.C2()
    B = w.default_batch  # 65,536
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, B)
    root = w.d["root_cluster"] + 1
    bt.assignfactors(params, tips)
    succ, iscal = bt.calibrate(None, 1)
    assert succ.all() and (bt.status() == 0).all()
    ll = bt.integratebelief(root, want_mu=False)[1]
    # (i) every element against the C twin of the oracle
    co = COracle.from_plan_dict(w.d).run_batch(params, tips, root_belief=w.d["root_cluster"], want_fe=True)
    assert (co["status"] == 0).all()
    assert np.max(np.abs(ll / co["loglik"] - 1)) <= TOL
    # (ii) calibrated clique tree: every belief integrates to the log-likelihood; factored energy too
    for j in range(1, plan.nclusters + plan.nsepsets + 1):
        if bt.dimension(j) > 0:
            assert np.max(np.abs(bt.integratebelief(j, want_mu=False)[1] / ll - 1)) <= 1e-9, j
    fe = bt.factored_energy()
    assert np.max(np.abs(fe[:, 2] / ll - 1)) <= 1e-9
    assert np.max(np.abs(fe[:, 2] / co["fe"][:, 2] - 1)) <= 1e-9
    # fixed point: a second calibration changes nothing beyond the residual tolerance
    b1 = [bt.get_belief(j) for j in (1, 7, 13)]
    succ, iscal = bt.calibrate(None, 1)
    assert succ.all() and iscal.all()
    for j, (J1, h1, g1) in zip((1, 7, 13), b1):
        J2, h2, g2 = bt.get_belief(j)
        assert relerr(J2, J1) <= 1e-9 and relerr(h2, h1) <= 1e-9 and relerr(g2, g1) <= 1e-9
    # determinism: reset + calibrate again is bit-identical
    bt.init_beliefs_reset_fromfactors()
    bt.calibrate(None, 1)
    ll2 = bt.integratebelief(root, want_mu=False)[1]
    assert np.array_equal(ll, ll2)


def test_c2_ragged_batch_sizes():
    # batch sizes that are not multiples of the warp / block / row-pitch granularity
    lib = get_lib("cuda")
    w = bench.C2()
    params, tips = w.inputs(1000, 0)
    co = COracle.from_plan_dict(w.d)
    plan = plan_of(w, lib)
    for B in (1, 31, 33, 127, 129, 1000):
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B)
        bt.assignfactors(params, tips[:B])
        succ, _ = bt.calibrate(None, 1)
        ll = bt.integratebelief(w.d["root_cluster"] + 1, want_mu=False)[1]
        ref = co.run_batch(params, tips[:B], root_belief=w.d["root_cluster"])["loglik"]
        assert succ.all() and np.max(np.abs(ll / ref - 1)) <= TOL, B


def test_c4_full_network_theta_grid_vs_cport():
    # the 10,000-tip network of BASELINE configs[3] at a reduced theta grid (256 of 4,096: the C twin
    # needs ~1 s for these); medium message shapes (8,8) and (16,8) through the shared-memory kernel
    lib = get_lib("cuda")
    w = bench.C4()
    B = 256
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    root = w.d["root_cluster"] + 1
    co = COracle.from_plan_dict(w.d)
    ref = co.run_batch(params, tips, ncolors=w.ncolors, root_belief=w.d["root_cluster"], B=B, **w.cpu_kw)
    assert (ref["status"] == 0).all()
    out = {}
    for mode in (-1, 8, 0):  # shared-memory kernel, cooperative kernel, thread-local generic kernel
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, factors=False, residuals=False)
        bt.set_coop_mode(mode)
        bt.assignfactors(params, tips, ncolors=w.ncolors)
        succ = bt.propagate_1traversal_postorder(0, update_residualnorm=False)
        assert succ.all()
        out[mode] = bt.integratebelief(root, want_mu=False)[1]
        assert np.max(np.abs(out[mode] / ref["loglik"] - 1)) <= TOL, mode
    assert np.array_equal(out[-1], out[8]) and np.array_equal(out[-1], out[0])  # bit-identical variants
    # full calibration on the big tree: beliefs far from the root agree on the log-likelihood
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, 64, factors=False, residuals=True)
    bt.assignfactors(params[:64], tips, ncolors=w.ncolors)
    succ, _ = bt.calibrate(None, 1)
    assert succ.all()
    ll = bt.integratebelief(root, want_mu=False)[1]
    rng = np.random.default_rng(5)
    for j in rng.choice(plan.nclusters, size=12, replace=False) + 1:
        if bt.dimension(int(j)) > 0:
            assert np.max(np.abs(bt.integratebelief(int(j), want_mu=False)[1] / ll - 1)) <= 1e-8, j


def test_pipelined_calibration_is_bit_identical():
    # calibrate in element chunks on several streams (pgbp_batch_set_pipeline) == one stream
    lib = get_lib("cuda")
    w = bench.C2()
    B = 20000  # chunks of 128-element granularity, last one ragged
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    out = {}
    for nchunks in (1, 3, 4, -1):
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B)
        bt.set_pipeline(nchunks)
        bt.assignfactors(params, tips)
        n0 = bt.launch_count(reset=True)
        succ, iscal = bt.calibrate(None, 2, auto=True)
        nl = bt.launch_count()
        ll = bt.integratebelief(w.d["root_cluster"] + 1, want_mu=False)[1]
        out[nchunks] = (succ, iscal, ll, bt.get_belief(7), bt.get_residual(plan.nclusters + 1, plan.sepset_clusters[0][0] + 1), nl)
    assert out[3][5] == 3 * out[1][5] and out[4][5] == 4 * out[1][5]
    for k in (3, 4, -1):
        assert np.array_equal(out[1][0], out[k][0]) and np.array_equal(out[1][1], out[k][1])
        assert np.array_equal(out[1][2], out[k][2])
        for x, y in zip(out[1][3], out[k][3]):
            assert np.array_equal(x, y)
        for x, y in zip(out[1][4][:3], out[k][4][:3]):
            assert np.array_equal(x, y)
    assert out[1][0].all() and out[1][1].all()


def test_c3_muller_loopy_bethe_vs_cport():
    # BASELINE configs[2]: muller_2022, Bethe cluster graph, regularizebeliefs_bycluster!, 10 iterations
    # over both spanning trees (62,240 messages per replicate), against the C twin on a sample of the
    # batch; CUDA-graph replay (third call) must reproduce the eager result bit for bit
    lib = get_lib("cuda")
    w = bench.C3()
    B = 1024
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    ref = COracle.from_plan_dict(w.d).run_batch(params, tips[:128], root_belief=w.d["root_cluster"], want_fe=True, **w.cpu_kw)
    assert (ref["status"] == 0).all()
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, B)
    bt.assignfactors(params, tips)
    fes = []
    for rep in range(3):  # eager, capture, replay
        bt.init_beliefs_reset_fromfactors()
        bt.init_messagecalibrationflags_reset()
        bt.regularizebeliefs_bycluster()
        succ, iscal = bt.calibrate(None, w.niter)
        assert succ.all()
        fes.append(bt.factored_energy())
        assert np.array_equal(iscal[:128], ref["iscal"])
    assert np.array_equal(fes[0], fes[1]) and np.array_equal(fes[0], fes[2])
    # the automatic strategy here is the tile-walk kernel (runs of narrow steps in one launch, wide steps on
    # their own); per-step launches and other lane counts must give the same bits
    n_auto = bt.launch_count(reset=True)
    for mode, lanes, wide in ((0, 0, 0), (1, 16, 16), (1, 4, 1000)):
        bt.set_tilewalk_mode(mode)
        if lanes:
            bt.set_tilewalk_params(lanes, wide)
        bt.init_beliefs_reset_fromfactors()
        bt.init_messagecalibrationflags_reset()
        bt.regularizebeliefs_bycluster()
        succ, iscal = bt.calibrate(None, w.niter)
        assert succ.all()
        assert np.array_equal(bt.factored_energy(), fes[0]), (mode, lanes, wide)
        if mode == 0:
            assert bt.launch_count(reset=True) > 10 * n_auto / 3
    # Tolerance (adjudicated, tests/test_adjudication.py): regularizebeliefs_bycluster! gives the
    # 800 factor-less variable clusters eps = 2.2e-16 (src/clustergraphbeliefs.jl:244) and the reference's recursion
    # evaluated in binary64 is unstable on them -- its OWN formulation (the C twin) ends 3e-5 away from the exact
    # (binary128) factored energy after 10 iterations, this library's fused formulation 4.5e-5, while the exact
    # answer itself moves by 1e-16 under 1-ulp input perturbations.  Stated tolerance of configs[2]-Bethe: 1e-4
    # against exact; against the binary64 twin the same bound.  In reference-order mode (PGBP_CAL_REFORDER) the
    # beliefs equal the twin bit for bit (next test).
    exact = COracle.from_plan_dict(w.d).run_batch(params, tips[:32], root_belief=w.d["root_cluster"], want_fe=True, quad=True,
                                                  **w.cpu_kw)["fe"]
    assert np.max(np.abs(fes[0][:32, 2] / exact[:, 2] - 1)) <= 1e-4
    assert np.max(np.abs(ref["fe"][:32, 2] / exact[:, 2] - 1)) <= 1e-4
    assert np.max(np.abs(fes[0][:128, 2] / ref["fe"][:, 2] - 1)) <= 1e-4
    assert np.max(np.abs(fes[0][:128, 0] / ref["fe"][:, 0] - 1)) <= 1e-4
    assert np.max(np.abs(fes[0][:128, 1] / ref["fe"][:, 1] - 1)) <= 1e-6
    # reference-order validation mode on the GPU: factored energy equal to the twin's to the clique-tree tolerance
    bt.set_tilewalk_mode(-1)
    bt.init_beliefs_reset_fromfactors()
    bt.init_messagecalibrationflags_reset()
    bt.regularizebeliefs_bycluster()
    succ, iscal = bt.calibrate(None, w.niter, reference_order=True)
    assert succ.all() and np.array_equal(iscal[:128], ref["iscal"])
    assert np.max(np.abs(bt.factored_energy()[:128] / ref["fe"] - 1)) <= TOL


def test_c3_muller_loopy_ltrip_vs_cport():
    # BASELINE configs[2], LTRIP(net) cluster graph (801 clusters / 1158 sepsets, sepsets of up to two nodes): the
    # automatic strategy is the tile-walk kernel; per-step launches must give the same bits, the C twin the same
    # calibration flags and factored energies to 1e-10
    lib = get_lib("cuda")
    w = bench.C3L()
    B = 512
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    ref = COracle.from_plan_dict(w.d).run_batch(params, tips[:64], root_belief=w.d["root_cluster"], want_fe=True, **w.cpu_kw)
    assert (ref["status"] == 0).all()
    out = {}
    for mode in (-1, 0):
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B)
        bt.set_tilewalk_mode(mode)
        bt.assignfactors(params, tips)
        bt.regularizebeliefs_bycluster()
        bt.launch_count(reset=True)
        succ, iscal = bt.calibrate(None, w.niter)
        assert succ.all()
        out[mode] = (iscal, bt.factored_energy(), bt.launch_count())
    assert np.array_equal(out[-1][0], out[0][0]) and np.array_equal(out[-1][1], out[0][1])
    assert out[-1][2] * 10 < out[0][2]
    assert np.array_equal(out[-1][0][:64], ref["iscal"])
    for k in range(3):  # the LTRIP graph has no factor-less clusters: the clique-tree tolerance holds
        assert np.max(np.abs(out[-1][1][:64, k] / ref["fe"][:, k] - 1)) <= TOL


def test_c5_shapes_reduced_network_vs_cport():
    # BASELINE configs[4] at reduced size (2,000 tips instead of 100,000; the full network is a bench
    # workload: python bench.py --workload c5): MvFullBM p = 16, sender dimensions 16 / 32 / 48 -- the
    # (16,16) messages through the shared-memory kernel, (32,16) through the generic kernel --
    # full calibration, log-likelihood at the root and at distant clusters against the C twin
    lib = get_lib("cuda")
    w = bench.C5(ntips=2000, nretic=200)
    B = 6
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    root = w.d["root_cluster"] + 1
    ref = COracle.from_plan_dict(w.d).run_batch(params, tips, root_belief=w.d["root_cluster"], B=B, **w.cpu_kw)
    assert (ref["status"] == 0).all()
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, factors=False, residuals=False)
    bt.assignfactors(params, tips)
    succ, _ = bt.calibrate(None, 1, update_residualnorm=False)
    assert succ.all()
    ll = bt.integratebelief(root, want_mu=False)[1]
    assert np.max(np.abs(ll / ref["loglik"] - 1)) <= TOL
    rng = np.random.default_rng(6)
    for j in rng.choice(plan.nclusters, size=10, replace=False) + 1:
        if bt.dimension(int(j)) > 0:
            assert np.max(np.abs(bt.integratebelief(int(j), want_mu=False)[1] / ll - 1)) <= 1e-8, j


def test_c5_full_network_vs_cport():
    # BASELINE configs[4] at FULL size: the 100,000-tip network with 10,000 reticulations (219,999 nodes, clique tree of
    # 209,998 clusters, sender dimensions 16 / 32 / 48), MvFullBM p = 16, 8 replicates (7.4 GB of state): full
    # calibration (419,994 messages per replicate), log-likelihood at the root against the C twin, agreement of
    # distant clusters on the log-likelihood (calibrated clique tree), determinism of a repeated run
    lib = get_lib("cuda")
    w = bench.C5()
    B = 8
    params, tips = w.inputs(B, 0)
    plan = plan_of(w, lib)
    root = w.d["root_cluster"] + 1
    ref = COracle.from_plan_dict(w.d).run_batch(params, tips, root_belief=w.d["root_cluster"], B=B, **w.cpu_kw)
    assert (ref["status"] == 0).all()
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, factors=False, residuals=False)
    bt.assignfactors(params, tips)
    succ, _ = bt.calibrate(None, 1, update_residualnorm=False)
    assert succ.all() and (bt.status() == 0).all()
    ll = bt.integratebelief(root, want_mu=False)[1]
    assert np.max(np.abs(ll / ref["loglik"] - 1)) <= TOL
    rng = np.random.default_rng(7)
    for j in rng.choice(plan.nclusters + plan.nsepsets, size=16, replace=False) + 1:
        if bt.dimension(int(j)) > 0:
            assert np.max(np.abs(bt.integratebelief(int(j), want_mu=False)[1] / ll - 1)) <= 1e-9, j
    bt.assignfactors(params, tips)
    bt.calibrate(None, 1, update_residualnorm=False)
    assert np.array_equal(bt.integratebelief(root, want_mu=False)[1], ll)
