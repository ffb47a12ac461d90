"""Parity of the CUDA path (through the C ABI) against the oracle and against
the reference's golden values.

Every test runs twice: `emul` (host emulation of the kernel bodies; CPU; checks
the host logic -- plan compiler, index tables, ABI) and `cuda` (marked gpu; the
product).  Tolerance: north_star asks for <= 1e-10 relative on beliefs and
log-likelihoods; the assertions below use TOL = 1e-10 (observed ~1e-15).
Schedules and scope indices are compared bit-exactly.
"""
import json
import math
import os

import numpy as np
import pytest

import pgbp_b200
from harness import BACKENDS, Case, get_lib, relerr
from oracle import beliefs as OB
from oracle import bp as OBP
from oracle import clustergraph as CG
from oracle import models as M

TOL = 1e-10
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
NAN = math.nan
TBL = np.array([[10, 1.0], [10, 0.9], [NAN, 1.0], [0, -1.0]])
TAXA = ["A", "B1", "B2", "C"]
HINT9 = GOLD["canonicalform_beliefnodelabels"][:7]


def check_all_beliefs(case, batch, cgbs, tol=TOL, elems=None, floor=0.0):
    """floor > 0: quantities that are exact zeros in exact arithmetic (a missing tip trait marginalised out of
    its edge factor: j - j*j/j) are roundoff of either sign on both sides; differences are then measured
    against max(|reference|, floor) instead of |reference| alone."""
    elems = range(len(cgbs)) if elems is None else elems
    worst = 0.0

    def rel(a, b):
        r = relerr(a, b)
        if floor and np.size(b):
            r = min(r, float(np.max(np.abs(np.asarray(a, float) - np.asarray(b, float)))) / floor)
        return r
    for j in range(1, len(case.b) + 1):
        J, h, g = batch.get_belief(j)
        for e in elems:
            ob = cgbs[e].belief[j - 1]
            worst = max(worst, rel(J[e], ob.J), rel(h[e], ob.h), rel(g[e], ob.g))
    assert worst <= tol, worst
    return worst


# ------------------------------------------------------------------ plan / schedule (bit-exact)
def mirror_steps(nbeliefs, nclusters, frm, sep, to):
    """Independent restatement of the launch-step rule: a message
    runs one step after the last write of its sender and after the last access
    of its receiver / sepset."""
    lw, lr, steps = [-1] * nbeliefs, [-1] * nbeliefs, []
    for f, s, t in zip(frm, sep, to):
        st = max(lw[f] + 1, max(lw[t], lr[t]) + 1, max(lw[s], lr[s]) + 1)
        steps.append(st)
        lr[f] = max(lr[f], st)
        lw[t] = st
        lw[s] = st
    return steps


@pytest.mark.parametrize("backend", BACKENDS)
def test_plan_schedule_and_scope_bit_exact(backend):
    lib = get_lib(backend)
    tbl = np.array(GOLD["lazaridis_x"]).reshape(-1, 1).repeat(3, axis=1)
    net_taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    m = M.MvFullBrownianMotion(np.eye(3), np.zeros(3))
    case = Case(GOLD["lazaridis"], "cliquetree", tbl, net_taxa, m, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    assert case.sched[0][2] == GOLD["lazaridis_sched_parent"] and case.sched[0][3] == GOLD["lazaridis_sched_child"]
    # in-scope node counts with a fixed root, times p = 3
    assert case.plan.belief_dim[:17] == [3 * k for k in (3, 3, 1, 1, 0, 3, 4, 4, 1, 1, 1, 4, 4, 4, 4, 1, 1)]
    pa, ch = [j - 1 for j in case.sched[0][2]], [j - 1 for j in case.sched[0][3]]
    nb, nc = len(case.b), case.nclusters
    sepidx = {frozenset(b.metadata): j for j, b in enumerate(case.b) if j >= nc}
    labs = case.cg.labels
    for direction in (0, 1):
        lv = case.plan.levels(0, direction)
        n = len(pa)
        order = range(n - 1, -1, -1) if direction == 0 else range(n)
        frm = [(ch if direction == 0 else pa)[i] for i in order]
        to = [(pa if direction == 0 else ch)[i] for i in order]
        sep = [sepidx[frozenset((labs[f], labs[t]))] for f, t in zip(frm, to)]
        steps = mirror_steps(nb, nc, frm, sep, to)
        # library output is in execution order; index it back by reference position
        by_ref = {int(r): k for k, r in enumerate(lv["ref"])}
        assert sorted(by_ref) == list(range(n))
        for r in range(n):
            k = by_ref[r]
            assert (lv["frm"][k], lv["sepset"][k], lv["to"][k], lv["step"][k]) == (frm[r], sep[r], to[r], steps[r])
        assert lv["nsteps"] == max(steps) + 1
        assert list(lv["step"]) == sorted(lv["step"])
    # algorithmic bytes of one calibration (SURVEY 8d worked example): 49 232 B
    by = case.plan.traversal_cost(0, 0, True)[0] + case.plan.traversal_cost(0, 1, True)[0]
    assert by == 49232.0
    # scope maps == oracle's scopeindex
    for j in range(nc, nb):
        s = case.b[j]
        for lab in s.metadata:
            c = case.b[labs.index(lab)]
            assert list(pgbp_b200.scopeindex(s.nodelabel, s.inscope, c.nodelabel, c.inscope)) == list(OB.scopeindex(s, c))


@pytest.mark.parametrize("backend", BACKENDS)
def test_set_get_roundtrip_and_api_errors(backend):
    lib = get_lib(backend)
    case = Case(GOLD["netstr_named"], "cliquetree", TBL, TAXA, M.MvDiagBrownianMotion([2, 1], [3, -3], [0.1, 10]), lib,
                order_hint=HINT9, with_families=False)
    B = 37  # ragged: not a multiple of 32
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    rng = np.random.default_rng(1)
    for j in range(1, len(case.b) + 1):
        m = bt.dimension(j)
        A = rng.normal(size=(B, m, m))
        J = A + A.transpose(0, 2, 1)
        h, g = rng.normal(size=(B, m)), rng.normal(size=B)
        bt.set_belief(j, J, h, g)
        J2, h2, g2 = bt.get_belief(j)
        assert np.array_equal(J, J2) and np.array_equal(h, h2) and np.array_equal(g, g2)
    assert (bt.status() == 0).all()
    with pytest.raises(pgbp_b200.PgbpError):
        bt.assignfactors(pgbp_b200.bm_params([np.eye(2)], [0, 0]), np.zeros((1, 4, 2)))  # plan has no family table
    with pytest.raises(pgbp_b200.PgbpError):
        bt.propagate_belief(1, len(case.b), 2)  # that sepset does not join clusters 1 and 2
    with pytest.raises(pgbp_b200.PgbpError):
        bt.calibrate([5])
    bt0 = pgbp_b200.BatchedClusterGraphBelief(case.plan, 3, factors=False, residuals=False)
    with pytest.raises(pgbp_b200.PgbpError):
        bt0.factored_energy()
    with pytest.raises(pgbp_b200.PgbpError):
        bt0.calibrate()  # residual tracking needs the residual arrays
    assert bt0.calibrate(update_residualnorm=False)[0].all()


# ------------------------------------------------------------------ reference known answers
@pytest.mark.parametrize("backend", BACKENDS)
def test_canonicalform_six_messages(backend):
    # test/test_canonicalform.jl:65-109: six explicit propagate_belief! calls, then integratebelief!
    lib = get_lib(backend)
    m = M.UnivariateBrownianMotion(2, 3, 0)
    case = Case(GOLD["netstr_named"], "cliquetree", TBL[:, 1:], TAXA, m, lib, order_hint=HINT9)
    B = 3
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([2.0], [3.0]), TBL[None, :, 1:])
    e = {k.number: k for k in case.net.edges}
    J, h, g = bt.get_belief(1)
    np.testing.assert_allclose(J[0], 0.5 / e[4].length * np.array([[1, -1], [-1, 1]]), rtol=1e-14)
    assert (h == 0).all() and g[0] == pytest.approx(-math.log(2 * math.pi * e[4].length * 2) / 2, rel=1e-14)
    bpv = 0.5 / (e[7].gamma ** 2 * e[7].length + e[5].gamma ** 2 * e[5].length)
    J, h, g = bt.get_belief(4)
    np.testing.assert_allclose(J[1], bpv * np.array([[1, -.9, -.1], [-.9, .81, .09], [-.1, .09, .01]]), rtol=1e-13)
    for to, s, fr in [(1, 8, 2), (1, 9, 3), (4, 10, 1), (4, 12, 6), (4, 13, 7), (5, 11, 4)]:
        pgbp_b200.propagate_belief(bt, to, s, fr)
    mu, ll = pgbp_b200.integratebelief(bt, 5)
    assert (bt.status() == 0).all()
    assert np.all(np.abs(ll / GOLD["canonicalform_loglik"] - 1) < 1e-14)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case_", GOLD["evomodels"], ids=lambda c: c["id"])
def test_evomodels_postorder_loglik(backend, case_):
    # test/test_evomodels.jl:52-264; factors assigned by the host-side model code and uploaded
    # (covers OU, missing data and every root type), one postorder + integratebelief! on the device
    lib = get_lib(backend)
    cols = {"y": [1], "x": [0], "xy": [0, 1]}[case_["traits"]]
    model = getattr(M, case_["model"])(*eval(case_["args"], {"inf": math.inf, "np": np}))
    case = Case(GOLD["netstr_named"], "cliquetree", TBL[:, cols], TAXA, model, lib, schedule="spanningtree",
                with_families=False)
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 2)
    case.upload(bt, [case.oracle_cgb()] * 2)
    spt = case.sched[0]
    assert pgbp_b200.propagate_1traversal_postorder(bt, spt).all()
    _, ll = pgbp_b200.integratebelief(bt, spt[2][0])
    assert np.all(np.abs(ll / case_["loglik"] - 1) <= TOL)


@pytest.mark.parametrize("backend", BACKENDS)
def test_lazaridis_docs_numbers(backend):
    # docs/src/man/getting_started.md:184-189, 283-291
    lib = get_lib(backend)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    tbl = np.array(GOLD["lazaridis_x"]).reshape(-1, 1)
    case = Case(GOLD["lazaridis"], "cliquetree", tbl, taxa, M.UnivariateBrownianMotion(1, 0), lib,
                order_hint=GOLD["lazaridis_cluster_labels"])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 4)
    pgbp_b200.assignfactors(bt, pgbp_b200.bm_params([1.0], [0.0]), tbl[None])
    J, h, g = bt.get_belief(1)
    np.testing.assert_allclose(J[3], np.array(GOLD["lazaridis_b1_J"]), rtol=1e-13)
    assert g[3] == pytest.approx(GOLD["lazaridis_b1_g"], rel=1e-13)
    succ, iscal = pgbp_b200.calibrate(bt, case.sched)
    assert succ.all() and not iscal.any()
    _, norm = pgbp_b200.integratebelief(bt, 1)
    assert np.all(np.abs(norm / GOLD["lazaridis_norm"] - 1) < 1e-12)
    fe = pgbp_b200.factored_energy(bt)
    assert np.all(np.abs(fe[:, 2] / GOLD["lazaridis_fe"] - 1) < 1e-12)
    succ, iscal = pgbp_b200.calibrate(bt, case.sched)  # second pass: residuals are now ~0
    assert succ.all() and iscal.all()


@pytest.mark.parametrize("backend", BACKENDS)
def test_lazy_factor_snapshot_is_the_assigned_belief(backend):
    # ClusterFactor = copy of the cluster belief right after assignfactors! (src/clustergraphbeliefs.jl:106).
    # The library takes the snapshot lazily (K1 re-run into the factor array at the first read): a snapshot
    # read right away, one read after a calibration, and the belief read right after assignment must be
    # the same bits; a second assignment replaces a pending snapshot; init_factors_frombeliefs overrides it.
    lib = get_lib(backend)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    rng = np.random.default_rng(5)
    p, B = 2, 6
    R = np.array([[2.0, 0.5], [0.5, 1.0]])
    data = rng.normal(size=(2, B, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0, 0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    par = pgbp_b200.bm_params([R], np.zeros(p))
    nc = case.nclusters
    early = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    early.assignfactors(par, data[1])
    fac_early = [early.get_factor(j) for j in range(1, nc + 1)]  # materialised before anything else happens
    bel = [early.get_belief(j) for j in range(1, nc + 1)]
    late = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    late.assignfactors(par, data[0])  # pending snapshot, replaced by the next call
    late.assignfactors(par, data[1])
    succ, _ = late.calibrate(case.sched)
    assert succ.all()
    early.calibrate(case.sched)
    fe_late = late.factored_energy()  # first reader of the factors
    assert np.array_equal(fe_late, early.factored_energy())
    for j in range(nc):
        for x, y, z in zip(fac_early[j], late.get_factor(j + 1), bel[j]):
            assert np.array_equal(x, y) and np.array_equal(x, z)
    late.init_beliefs_reset_fromfactors()
    for j in range(nc):
        for x, y in zip(late.get_belief(j + 1), bel[j]):
            assert np.array_equal(x, y)
    late.assignfactors(par, data[0])  # pending again ...
    late.calibrate(case.sched)
    late.init_factors_frombeliefs()  # ... and overridden by an explicit snapshot of the calibrated beliefs
    for j in range(nc):
        for x, y in zip(late.get_factor(j + 1), late.get_belief(j + 1)):
            assert np.array_equal(x, y)


# ------------------------------------------------------------------ device factor assignment
ASSIGN_CASES = [
    ("uni_fixed", lambda: M.UnivariateBrownianMotion(2, 3, 0), [1], lambda: ([[[2.0]]], [3.0], None)),
    ("uni_improper", lambda: M.UnivariateBrownianMotion(2, 3, math.inf), [1], lambda: ([[[2.0]]], [3.0], [[math.inf]])),
    ("uni_random", lambda: M.UnivariateBrownianMotion(2, 3, 0.4), [1], lambda: ([[[2.0]]], [3.0], [[0.4]])),
    ("full_fixed", lambda: M.MvFullBrownianMotion([[2.0, 0.5], [0.5, 1.0]], [3.0, -3.0]), [0, 1],
     lambda: ([[[2.0, 0.5], [0.5, 1.0]]], [3.0, -3.0], None)),
    ("full_random", lambda: M.MvFullBrownianMotion([[2.0, 0.5], [0.5, 1.0]], [3.0, -3.0], [[0.1, 0.01], [0.01, 0.2]]), [0, 1],
     lambda: ([[[2.0, 0.5], [0.5, 1.0]]], [3.0, -3.0], [[0.1, 0.01], [0.01, 0.2]])),
    ("full_improper", lambda: M.MvFullBrownianMotion([[2.0, 0.5], [0.5, 1.0]], [3.0, -3.0], [[math.inf, 0], [0, math.inf]]), [0, 1],
     lambda: ([[[2.0, 0.5], [0.5, 1.0]]], [3.0, -3.0], [[math.inf, 0], [0, math.inf]])),
    ("diag_random", lambda: M.MvDiagBrownianMotion([2, 1], [3, -3], [0.1, 10]), [0, 1],
     lambda: ([[2.0, 1.0]], [3.0, -3.0], [0.1, 10.0])),
]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name,mk,cols,par", ASSIGN_CASES, ids=[c[0] for c in ASSIGN_CASES])
def test_assignfactors_device_vs_oracle(backend, name, mk, cols, par):
    lib = get_lib(backend)
    tbl = TBL.copy()
    tbl[2, 0] = 7.5  # device assignment: no missing data
    tbl = tbl[:, cols]
    model = mk()
    case = Case(GOLD["netstr_named"], "cliquetree", tbl, TAXA, model, lib, schedule="spanningtree")
    B = 5
    rng = np.random.default_rng(3)
    data = tbl[None] + rng.normal(size=(B,) + tbl.shape)
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    rates, mu, v = par()
    bt.assignfactors(pgbp_b200.bm_params(rates, mu, v), data)
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
    check_all_beliefs(case, bt, cgbs)
    for j in range(1, case.nclusters + 1):  # factor snapshot
        J, h, g = bt.get_factor(j)
        assert relerr(J[1], cgbs[1].factor[j - 1].J) <= TOL and relerr(g[1], cgbs[1].factor[j - 1].g) <= TOL
    spt = case.sched[0]
    assert bt.propagate_1traversal_postorder(spt).all()
    _, ll = bt.integratebelief(spt[2][0])
    for e in range(B):
        OBP.propagate_1traversal_postorder(cgbs[e], *spt)
        assert abs(ll[e] / OBP.integratebelief_cgb(cgbs[e], spt[2][0])[1] - 1) <= TOL


MISSING_CASES = [c for c in GOLD["evomodels"] if c["traits"] in ("x", "xy") and "BrownianMotion" in c["model"]]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case_", MISSING_CASES, ids=lambda c: c["id"])
def test_assignfactors_device_missing_data_goldens(backend, case_):
    # trait-level scopes on the device: trait x is missing at tip B2, so the scoped K1 path
    # runs absorbleaf! + the marginalisations of src/beliefs.jl:833-857 per family.  Beliefs against the
    # oracle's assignfactors!, log-likelihood against the reference's goldens (test/test_evomodels.jl:107-262);
    # other replicates of the batch carry different values at the observed entries.
    lib = get_lib(backend)
    cols = {"x": [0], "xy": [0, 1]}[case_["traits"]]
    tbl = TBL[:, cols]
    args = eval(case_["args"], {"inf": math.inf, "np": np})
    model = getattr(M, case_["model"])(*args)
    edge_color = None
    if case_["model"] == "HeterogeneousBrownianMotion":
        colors = args[1] or {}
        edge_color = lambda num: colors.get(num, 1) - 1  # noqa: E731
        rates, mu, v = [np.array(r) for r in args[0]], args[2], (args[3] if len(args) > 3 else None)
    elif case_["model"] == "UnivariateBrownianMotion":
        rates, mu, v = [[[float(args[0])]]], [float(args[1])], [[float(args[2])]]
    elif case_["model"] == "MvDiagBrownianMotion":
        rates, mu, v = [np.array(args[0], float)], args[1], np.array(args[2], float)
    else:
        rates, mu, v = [np.array(args[0])], args[1], (np.array(args[2]) if len(args) > 2 else None)
    case = Case(GOLD["netstr_named"], "cliquetree", tbl, TAXA, model, lib, schedule="spanningtree", edge_color=edge_color)
    assert case.plan.families.get("mem_tpos") is not None
    B = 4
    rng = np.random.default_rng(17)
    data = np.repeat(tbl[None], B, axis=0)
    data[1:] += rng.normal(size=(B - 1,) + tbl.shape)  # NaN stays NaN
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params(rates, mu, v), data, ncolors=len(rates))
    assert (bt.status() == 0).all()
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
    check_all_beliefs(case, bt, cgbs, floor=1e-3)
    spt = case.sched[0]
    assert bt.propagate_1traversal_postorder(spt).all()
    _, ll = bt.integratebelief(spt[2][0])
    assert abs(ll[0] / case_["loglik"] - 1) <= TOL
    for e in range(B):
        OBP.propagate_1traversal_postorder(cgbs[e], *spt)
        assert abs(ll[e] / OBP.integratebelief_cgb(cgbs[e], spt[2][0])[1] - 1) <= TOL
    # a NaN where the plan expects a value is a per-element status, not a silent NaN
    bad = data.copy()
    bad[2, 0, 0] = NAN
    bt.assignfactors(pgbp_b200.bm_params(rates, mu, v), bad, ncolors=len(rates))
    stt = bt.status()
    assert stt[2] != 0 and (np.delete(stt, 2) == 0).all()


@pytest.mark.parametrize("backend", BACKENDS)
def test_assignfactors_device_missing_data_tree_and_joingraph(backend):
    # test/test_calibration.jl:107-129: tree, trait 2 observed at one tip only (internal nodes lose that trait
    # from their scope); every belief integrates to -7.578343735986344 after calibration.
    lib = get_lib(backend)
    tbl = np.array([[1, NAN], [1, NAN], [1, NAN], [1, 1.0]])
    m = M.MvDiagBrownianMotion([1, 1], [0, 0])
    case = Case("(((A:1.0, B:1.0)E:1.0, C:2.0)F:1.0, D:3.0)G;", "cliquetree", tbl, ["A", "B", "C", "D"], m, lib,
                schedule="spanningtree")
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 3)
    bt.assignfactors(pgbp_b200.bm_params([np.array([1.0, 1.0])], [0, 0]), tbl[None])
    cgb = case.oracle_cgb()
    check_all_beliefs(case, bt, [cgb] * 3, floor=1e-3)
    succ, _ = bt.calibrate(case.sched)
    assert succ.all()
    for j in range(1, len(case.b) + 1):
        assert np.all(np.abs(bt.integratebelief(j, want_mu=False)[1] / -7.578343735986344 - 1) <= TOL)
    # test/test_calibration.jl:131-185: level-3 network, join-graph structuring, improper root, one value missing
    tbl = np.array([[2.11, 30.0], [2.15, NAN]])
    R = [[1, 0.5], [0.5, 1]]
    m = M.MvFullBrownianMotion(R, [0, 0], [[math.inf, 0], [0, math.inf]])
    case = Case(GOLD["netstr_level3"], "jgs", tbl, ["A", "B"], m, lib, maxclustersize=3)
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 2)
    bt.assignfactors(pgbp_b200.bm_params([np.array(R, float)], [0, 0], [[math.inf, 0], [0, math.inf]]), tbl[None])
    assert (bt.status() == 0).all()
    check_all_beliefs(case, bt, [case.oracle_cgb()] * 2, floor=1e-3)


@pytest.mark.parametrize("backend", BACKENDS)
def test_assignfactors_heterogeneous_theta_grid(backend):
    # config-4 shape in miniature: painted rates (one hybrid with parents of different colours),
    # a grid of parameter vectors x one data set (product pairing)
    lib = get_lib(backend)
    tbl = TBL.copy()
    tbl[2, 0] = 7.5
    colors = {9: 2, 7: 2, 8: 2, 1: 3}
    rng = np.random.default_rng(7)
    NP = 6

    def rnd_spd(p):
        A = rng.normal(size=(p, p))
        return A @ A.T / p + 0.1 * np.eye(p)

    thetas = [[rnd_spd(2) for _ in range(3)] for _ in range(NP)]
    mus = rng.normal(size=(NP, 2))
    model0 = M.HeterogeneousBrownianMotion(thetas[0], colors, mus[0])
    case = Case(GOLD["netstr_named"], "cliquetree", tbl, TAXA, model0, lib, schedule="spanningtree",
                edge_color=lambda num: colors.get(num, 1) - 1)
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, NP)
    params = np.stack([pgbp_b200.bm_params(thetas[k], mus[k]) for k in range(NP)])
    bt.assignfactors(params, tbl[None], ncolors=3, pairing="product")
    cgbs = [case.oracle_cgb(model=M.HeterogeneousBrownianMotion(thetas[k], colors, mus[k])) for k in range(NP)]
    check_all_beliefs(case, bt, cgbs)
    spt = case.sched[0]
    assert bt.propagate_1traversal_postorder(spt, update_residualnorm=False).all()
    _, ll = bt.integratebelief(spt[2][0], want_mu=False)
    for k in range(NP):
        OBP.propagate_1traversal_postorder(cgbs[k], *spt)
        assert abs(ll[k] / OBP.integratebelief_cgb(cgbs[k], spt[2][0])[1] - 1) <= TOL
    # a non-PD rate matrix fails only its own element
    params[2, :4] = [1.0, 2.0, 2.0, 1.0]
    bt.assignfactors(params, tbl[None], ncolors=3, pairing="product")
    st = bt.status()
    assert st[2] != 0 and (np.delete(st, 2) == 0).all()


# ------------------------------------------------------------------ calibration
@pytest.mark.parametrize("backend", BACKENDS)
def test_calibrate_cliquetree_all_beliefs(backend):
    # test/test_calibration.jl:36-78
    lib = get_lib(backend)
    tbl_y = np.array([[1.0], [.9], [1], [-1]])
    m = M.UnivariateBrownianMotion(0.471474, 0, math.inf)
    case = Case(GOLD["netstr_named"], "cliquetree", tbl_y, TAXA, m, lib, schedule="spanningtree")
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 2)
    bt.assignfactors(pgbp_b200.bm_params([0.471474], [0.0], [math.inf]), tbl_y[None])
    pgbp_b200.calibrate(bt, case.sched)
    ll = -4.877930583154144
    for j in range(1, len(case.b) + 1):
        assert np.all(np.abs(bt.integratebelief(j)[1] / ll - 1) <= TOL)
    assert np.all(np.abs(pgbp_b200.factored_energy(bt)[:, 2] / ll - 1) <= TOL)
    root_ind = next(i for i, be in enumerate(case.b) if 1 in be.nodelabel) + 1
    mu, _ = bt.integratebelief(root_ind)
    assert mu[0, -1] == pytest.approx(-0.26000871507162693, rel=TOL)
    cgb = case.oracle_cgb()
    OBP.calibrate(cgb, case.sched)
    check_all_beliefs(case, bt, [cgb, cgb])
    # residuals and flags of every directed message
    labs = case.cg.labels
    for j in range(case.nclusters, len(case.b)):
        for lab_to in case.b[j].metadata:
            lab_from = [l for l in case.b[j].metadata if l != lab_to][0]
            dJ, dh, fl, _ = bt.get_residual(j + 1, labs.index(lab_to) + 1)
            r = cgb.messageresidual[(lab_to, lab_from)]
            assert relerr(dJ[0], r.dJ) <= TOL and relerr(dh[0], r.dh) <= TOL and bool(fl[0]) == r.iscalibrated_resid
    # the graph invariant survives both regularisations (test/test_calibration.jl:65-77)
    for reg in ("bynodesubtree", "bycluster"):
        bt.init_beliefs_reset_fromfactors()
        cgb2 = case.oracle_cgb()
        if reg == "bycluster":
            bt.regularizebeliefs_bycluster()
            OBP.regularizebeliefs_bycluster(cgb2, case.cg)
        else:
            bt.regularizebeliefs_bynodesubtree(OBP.bynodesubtree_program(cgb2, case.cg))
            OBP.regularizebeliefs_bynodesubtree(cgb2, case.cg)
        check_all_beliefs(case, bt, [cgb2, cgb2])
        pgbp_b200.calibrate(bt, case.sched)
        assert np.all(np.abs(bt.integratebelief(1)[1] / ll - 1) <= TOL)


@pytest.mark.parametrize("backend", BACKENDS)
def test_loopy_bethe_onschedule_autostop(backend):
    # test/test_calibration.jl:79-105: "calibration reached: iteration 5, schedule tree 1"
    lib = get_lib(backend)
    tbl_y = np.array([[-1.81358], [0.468158], [0.658486], [0.643821]])
    taxa = ["A", "B", "C", "D"]
    m = M.UnivariateBrownianMotion(0.0861249, 0)
    case = Case(GOLD["netstr_unnamed"], "bethe", tbl_y, taxa, m, lib)
    B = 4
    rng = np.random.default_rng(11)
    data = np.repeat(tbl_y[None], B, axis=0)
    data[1:] += 0.3 * rng.normal(size=(B - 1, 4, 1))
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([0.0861249], [0.0]), data)
    bt.regularizebeliefs_onschedule()
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
    for c in cgbs:
        OBP.regularizebeliefs_onschedule(c, case.cg)
    check_all_beliefs(case, bt, cgbs)
    succ, iscal, it = pgbp_b200.calibrate(bt, case.sched, 20, auto=True, info=True)
    assert succ.all() and iscal.all()
    assert tuple(it[0]) == (5, 1)
    infos = []
    for c in cgbs:
        assert all(OBP.calibrate(c, case.sched, 20, auto=True))
        infos.append(OBP.calibrate.last_info)
    assert [tuple(x) for x in it] == infos  # per-element auto-stop point, bit-exact
    check_all_beliefs(case, bt, cgbs, tol=TOL)
    ind = case.cg.labels.index("I3") + 1
    # the golden is the EXACT posterior root mean (a linear-model fit, test/test_calibration.jl:92); loopy BP stopped at
    # its 1e-5 residual tolerance approximates it to rtol 1e-5, the reference's own bound (test_calibration.jl:105) --
    # the device result equals the oracle's loopy result to 1e-10 (check_all_beliefs above)
    assert bt.integratebelief(ind)[0][0, -1] == pytest.approx(0.21511454631828986, rel=1e-5)
    fe = bt.factored_energy()
    for e in range(B):
        assert relerr(fe[e], np.array(OBP.factored_energy(cgbs[e]))) <= TOL


@pytest.mark.parametrize("backend", BACKENDS)
def test_joingraph_missing_data_bynodesubtree(backend):
    # test/test_calibration.jl:131-185: level-3 network, missing trait, improper root,
    # JoinGraphStructuring(3), one schedule tree per variable
    lib = get_lib(backend)
    tbl = np.array([[2.11, 30.0], [2.15, NAN]])
    m = M.MvFullBrownianMotion([[1, 0.5], [0.5, 1]], [0, 0], [[math.inf, 0], [0, math.inf]])

    def sched(case):
        out = []
        for n in case.net.vec_node:
            st = CG.nodesubtree_clusterlist(case.cg, n.name)
            if st[0]:
                out.append(st)
        return out

    case = Case(GOLD["netstr_level3"], "jgs", tbl, ["A", "B"], m, lib, schedule=sched, with_families=False, maxclustersize=3)
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 2)
    cgb = case.oracle_cgb()
    case.upload(bt, [cgb, cgb])
    bt.init_factors_frombeliefs()
    bt.regularizebeliefs_bynodesubtree(OBP.bynodesubtree_program(cgb, case.cg))
    OBP.regularizebeliefs_bynodesubtree(cgb, case.cg)
    succ, iscal, it = bt.calibrate(case.sched, 10, auto=True, info=True)
    assert succ.all() and iscal.all()
    assert all(OBP.calibrate(cgb, case.sched, 10, auto=True))
    assert tuple(it[0]) == OBP.calibrate.last_info
    i6 = case.cg.labels.index("I1I2I3") + 1
    mu, nrm = bt.integratebelief(i6)
    assert nrm[0] == pytest.approx(-1.390595772423, rel=TOL)
    np.testing.assert_allclose(mu[0], [2.121105154896223, 30.005552577448075, 2.1360649504455984,
                                       30.013032475222563, 2.128585052670943, 30.00929252633547], rtol=TOL)
    check_all_beliefs(case, bt, [cgb, cgb], tol=TOL)


# ------------------------------------------------------------------ failure semantics
@pytest.mark.parametrize("backend", BACKENDS)
def test_failed_cholesky_is_per_element_status(backend):
    # src/beliefupdates.jl:68-76, 640-644: the exception is returned, not thrown; here: status word
    lib = get_lib(backend)
    m = M.UnivariateBrownianMotion(2, 3, 0)
    case = Case(GOLD["netstr_named"], "cliquetree", TBL[:, 1:], TAXA, m, lib, order_hint=HINT9, schedule="spanningtree")
    B = 6
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    cgb = case.oracle_cgb()
    case.upload(bt, [cgb] * B)
    # element 4: make cluster 1's precision indefinite in the variable integrated out by its first message
    J, h, g = bt.get_belief(1)
    J[4] = np.array([[-1.0, 0.3], [0.3, 2.0]])
    bt.set_belief(1, J, h, g)
    spt = case.sched[0]
    succ = bt.propagate_1traversal_postorder(spt)
    assert list(succ) == [True] * 4 + [False] + [True]
    st = bt.status()
    lv = case.plan.levels(0, 0)
    k = next(i for i in range(len(lv["ref"])) if lv["frm"][i] == 0)
    assert st[4] != 0 and (st[4] >> 8) - 1 == lv["ref"][k] and (st[4] & 0xff) == 1
    _, ll = bt.integratebelief(spt[2][0])
    ok = [0, 1, 2, 3, 5]
    assert np.all(np.abs(ll[ok] / GOLD["canonicalform_loglik"] - 1) < 1e-13)
    # oracle agrees on which message fails
    b2 = [x.copy() for x in cgb.belief]
    b2[0].J[:] = J[4]
    ex = None
    n = len(spt[0])
    for r, i in enumerate(range(n - 1, -1, -1)):
        ss = b2[cgb.sepsetindex(spt[0][i], spt[1][i]) - 1]
        ex = OBP.propagate_belief(b2[spt[2][i] - 1], ss, b2[spt[3][i] - 1])
        if ex is not None:
            assert r == (st[4] >> 8) - 1 and ex.info == (st[4] & 0xff)
            break
    assert ex is not None


# ------------------------------------------------------------------ shapes: generic kernel, larger traits
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("p", [3, 5, 8, 16])
def test_multivariate_shapes_vs_oracle(backend, p):
    # p=3: register-resident kernels (config 2 shapes); p=5: mixed; p=8, 16: generic kernel
    # (config 4 / 5 shapes: m up to 3p or 4p, s up to 3p)
    lib = get_lib(backend)
    rng = np.random.default_rng(100 + p)
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    mu = rng.normal(size=p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    B = 3
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, mu)
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([R], mu), data)
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
    check_all_beliefs(case, bt, cgbs)
    succ, _ = bt.calibrate(case.sched)
    assert succ.all()
    for c in cgbs:
        OBP.calibrate(c, case.sched)
    worst = check_all_beliefs(case, bt, cgbs, tol=TOL)
    _, ll = bt.integratebelief(case.sched[0][2][0])
    fe = bt.factored_energy()
    for e in range(B):
        ref = OBP.integratebelief_cgb(cgbs[e], case.sched[0][2][0])[1]
        assert abs(ll[e] / ref - 1) <= TOL and abs(fe[e, 2] / ref - 1) <= TOL


# ------------------------------------------------------------------ walk kernel == level-parallel launches
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_walk_kernel_matches_level_parallel_and_oracle(backend, p):
    # one launch per calibration (each thread walks all messages of its replicate) must give the
    # same beliefs, residuals, flags and failure status as one launch per step
    lib = get_lib(backend)
    rng = np.random.default_rng(200 + p)
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    B = 70
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    out = {}
    for mode in (0, 1):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
        bt.set_walk_mode(mode)
        bt.assignfactors(pgbp_b200.bm_params([R], np.zeros(p)), data)
        # element 5: indefinite precision in cluster 10 (a leaf cluster of the clique tree)
        J, h, g = bt.get_belief(10)
        J[5] = -np.eye(p)
        bt.set_belief(10, J, h, g)
        n0 = bt.launch_count(reset=True)
        succ, iscal = bt.calibrate(case.sched)
        nl = bt.launch_count()
        succ2, iscal2 = bt.calibrate(case.sched)
        out[mode] = dict(succ=succ, iscal2=iscal2, st=bt.status(), nl=nl,
                         beliefs=[bt.get_belief(j) for j in range(1, len(case.b) + 1)],
                         res=[bt.get_residual(case.nclusters + 1 + j, case.plan.sepset_clusters[j][s] + 1)
                              for j in range(case.plan.nsepsets) for s in (0, 1)])
    if p <= 3:
        assert out[1]["nl"] == 2 and out[0]["nl"] > 10  # walk kernel + iscal reduction
    else:  # 4 nodes x 4 traits = 16 > register-resident limit: falls back to level-parallel launches
        assert out[1]["nl"] == out[0]["nl"]
    ok = np.arange(B) != 5
    assert (out[0]["succ"] == ok).all() and (out[1]["succ"] == ok).all()
    assert np.array_equal(out[0]["st"], out[1]["st"]) and out[0]["st"][5] != 0
    assert out[0]["iscal2"][ok].all() and out[1]["iscal2"][ok].all()
    for (J0, h0, g0), (J1, h1, g1) in zip(out[0]["beliefs"], out[1]["beliefs"]):
        assert np.array_equal(J0[ok], J1[ok]) and np.array_equal(h0[ok], h1[ok]) and np.array_equal(g0[ok], g1[ok])
    for r0, r1 in zip(out[0]["res"], out[1]["res"]):
        assert np.array_equal(r0[0][ok], r1[0][ok]) and np.array_equal(r0[2][ok], r1[2][ok])
    for e in (0, 6, B - 1):
        cgb = case.oracle_cgb(tbl=data[e])
        OBP.calibrate(cgb, case.sched)
        OBP.calibrate(cgb, case.sched)
        for j, (J1, h1, g1) in enumerate(out[1]["beliefs"]):
            ob = cgb.belief[j]
            assert max(relerr(J1[e], ob.J), relerr(h1[e], ob.h), relerr(g1[e], ob.g)) <= TOL


# ------------------------------------------------------------------ cooperative (G lanes / element) kernel
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("p", [4, 5, 6, 7, 8, 12, 16])
def test_cooperative_kernel_matches_generic_and_oracle(backend, p):
    # medium shapes (sender dimension 13..32: lazaridis 3- and 4-node cliques at p = 4..8) run in
    # the cooperative kernel; it must agree BIT FOR BIT with the one-thread-per-element generic
    # kernel (same per-entry update order, explicit FMAs) in beliefs, residuals, flags and status,
    # and with the oracle to 1e-10.  B is not a multiple of the elements per warp / block.
    lib = get_lib(backend)
    rng = np.random.default_rng(300 + p)
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    mu = rng.normal(size=p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    B = 37
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, mu)
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    big = [j for j in range(1, case.nclusters + 1) if case.b[j - 1].dimension() > 12]
    assert big
    out = {}
    # modes: 0 generic, -1 automatic (shared-memory kernel; its multi-warp form for integrated dimensions
    # 12..16), 1 single-warp shared-memory kernel, 2 multi-warp form with 4 warps, 4 / 8 cooperative lanes
    for mode in (0, -1, 1, 2, 4, 8):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
        bt.set_coop_mode(mode)
        bt.assignfactors(pgbp_b200.bm_params([R], mu), data)
        # element 5: indefinite precision in a large cluster -> Cholesky failure in its first message;
        # element 7: a large cluster identically zero -> the "missing data" shortcut (message unchanged)
        J, h, g = bt.get_belief(big[0])
        J[5] = -np.eye(J.shape[1])
        bt.set_belief(big[0], J, h, g)
        J, h, g = bt.get_belief(big[-1])
        J[7] = 0.0
        h[7] = 0.0
        bt.set_belief(big[-1], J, h, g)
        succ, iscal = bt.calibrate(case.sched)
        succ2, iscal2 = bt.calibrate(case.sched)
        out[mode] = dict(succ=succ, iscal2=iscal2, st=bt.status(),
                         beliefs=[bt.get_belief(j) for j in range(1, len(case.b) + 1)],
                         res=[bt.get_residual(case.nclusters + 1 + j, case.plan.sepset_clusters[j][s] + 1)
                              for j in range(case.plan.nsepsets) for s in (0, 1)])
    ok = ~np.isin(np.arange(B), [5, 7])
    okz = np.arange(B) != 5
    for mode in (-1, 1, 2, 4, 8):
        assert np.array_equal(out[0]["st"], out[mode]["st"]) and out[0]["st"][5] != 0
        assert np.array_equal(out[0]["succ"], out[mode]["succ"]) and np.array_equal(out[0]["iscal2"], out[mode]["iscal2"])
        for (J0, h0, g0), (J1, h1, g1) in zip(out[0]["beliefs"], out[mode]["beliefs"]):
            assert np.array_equal(J0[okz], J1[okz], equal_nan=True) and np.array_equal(h0[okz], h1[okz], equal_nan=True)
            assert np.array_equal(g0[okz], g1[okz], equal_nan=True)
        for r0, r1 in zip(out[0]["res"], out[mode]["res"]):
            for x0, x1 in zip(r0[:3], r1[:3]):
                assert np.array_equal(x0[okz], x1[okz], equal_nan=True)
    assert out[-1]["succ"][ok].all() and out[-1]["iscal2"][ok].all()
    for e in (0, 6, B - 1):
        cgb = case.oracle_cgb(tbl=data[e])
        OBP.calibrate(cgb, case.sched)
        OBP.calibrate(cgb, case.sched)
        for j, (J1, h1, g1) in enumerate(out[-1]["beliefs"]):
            ob = cgb.belief[j]
            assert max(relerr(J1[e], ob.J), relerr(h1[e], ob.h), relerr(g1[e], ob.g)) <= TOL
    # element 7 against the oracle with the same zeroed cluster
    cgb = case.oracle_cgb(tbl=data[7])
    cgb.belief[big[-1] - 1].J[:] = 0.0
    cgb.belief[big[-1] - 1].h[:] = 0.0
    OBP.calibrate(cgb, case.sched, verbose=False)
    for j, (J1, h1, g1) in enumerate(out[-1]["beliefs"]):
        ob = cgb.belief[j]
        if np.all(np.isfinite(ob.J)) and np.all(np.isfinite(J1[7])):
            assert max(relerr(J1[7], ob.J), relerr(h1[7], ob.h)) <= TOL


@pytest.mark.parametrize("backend", BACKENDS)
def test_assignfactors_nan_tip_data_sets_status(backend):
    # device factor assignment needs complete tip data: an element whose table holds a NaN is flagged
    # (status PGBP_STATUS(0x7ffffa, trait)), the others are unaffected
    lib = get_lib(backend)
    p = 2
    R = np.array([[1.0, 0.3], [0.3, 2.0]])
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    rng = np.random.default_rng(9)
    data = rng.normal(size=(6, 7, p))
    data[4, 2, 1] = np.nan
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 6)
    bt.assignfactors(pgbp_b200.bm_params([R], np.zeros(p)), data)
    st = bt.status()
    assert st[4] != 0 and (st[4] >> 8) - 1 == 0x7ffffa and (st[4] & 0xff) == 2
    assert (np.delete(st, 4) == 0).all()
    succ, _ = bt.calibrate(case.sched)
    assert not succ[4] and np.delete(succ, 4).all()


# ------------------------------------------------------------------ residual KL divergence (update_residualkldiv)
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("p", [1, 3])
def test_residual_kldiv_matches_oracle(backend, p):
    # residual_kldiv! (src/beliefs.jl:1060-1075; oracle pinned by the golden 1.215973 of
    # test/test_calibration.jl:13-33): KL divergence of every directed message after 1 and 2
    # calibrations; -1 where a sepset belief is not yet proper (left untouched), 0 for empty sepsets
    lib = get_lib(backend)
    rng = np.random.default_rng(400 + p)
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    B = 4
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([R], np.zeros(p)), data)
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
    labs = case.cg.labels
    for rnd in range(2):
        succ, _ = bt.calibrate(case.sched, update_residualkldiv=True)
        assert succ.all()
        for c in cgbs:
            OBP.calibrate(c, case.sched, update_residualkldiv=True)
        for j in range(case.plan.nsepsets):
            a, b_ = case.plan.sepset_clusters[j]
            for to, frm in ((a, b_), (b_, a)):
                kl = bt.get_residual(case.nclusters + 1 + j, to + 1)[3]
                for e in range(B):
                    ref = cgbs[e].messageresidual[(labs[to], labs[frm])].kldiv
                    if ref in (-1.0, 0.0) or abs(ref) < 1e-12:
                        assert abs(kl[e] - ref) <= TOL, (rnd, j, to, e, kl[e], ref)
                    else:
                        assert abs(kl[e] / ref - 1) <= 1e-8, (rnd, j, to, e, kl[e], ref)
    # second calibration of a clique tree changes nothing: all KL divergences ~ 0
    kls = np.array([bt.get_residual(case.nclusters + 1 + j, case.plan.sepset_clusters[j][0] + 1)[3]
                    for j in range(case.plan.nsepsets)])
    assert np.all(np.abs(kls) <= 1e-5)


@pytest.mark.parametrize("backend", BACKENDS)
def test_tilewalk_kernel_matches_level_parallel(backend):
    # one launch per traversal (block barrier between steps) == one launch per step and shape, bit for bit,
    # on a loopy Bethe graph with univariate traits (sender dimension <= 3), incl. auto-stop and a failure
    lib = get_lib(backend)
    taxa = ["A", "B", "C", "D"]
    rng = np.random.default_rng(3)
    B = 45
    data = rng.normal(size=(B, 4, 1))
    model = M.UnivariateBrownianMotion(1.0, 0.0)
    case = Case(GOLD["netstr_unnamed"], "bethe", data[0], taxa, model, lib)
    out = {}
    # modes: 0 = per-step launches; 1 = tile-walk; 2 = tile-walk with steps wider than 1 message split off
    # into ordinary launches (the hybrid path of deep schedules with a few wide steps), 16 lanes
    for mode in (0, 1, 2):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
        bt.set_tilewalk_mode(min(mode, 1))
        if mode == 2:
            bt.set_tilewalk_params(lanes=16, wide=1)
        bt.assignfactors(pgbp_b200.bm_params([1.0], [0.0]), data)
        bt.regularizebeliefs_bycluster()
        jbig = next(j for j in range(1, case.nclusters + 1) if case.b[j - 1].dimension() >= 2)
        J, h, g = bt.get_belief(jbig)
        J[5] = -np.eye(J.shape[1])  # element 5: indefinite cluster -> failure status
        bt.set_belief(jbig, J, h, g)
        bt.launch_count(reset=True)
        succ, iscal, info = bt.calibrate(case.sched, 20, auto=True, info=True)
        nl = bt.launch_count()
        out[mode] = (succ, iscal, info, bt.status(), [bt.get_belief(j) for j in range(1, len(case.b) + 1)], nl,
                     bt.factored_energy())
    assert out[1][5] < out[2][5] < out[0][5]
    ok = np.arange(B) != 5
    assert out[0][3][5] != 0 and out[0][0][ok].all() and out[0][1][ok].all()
    for mode in (1, 2):
        for k in range(4):
            assert np.array_equal(out[0][k], out[mode][k])
        for (J0, h0, g0), (J1, h1, g1) in zip(out[0][4], out[mode][4]):
            assert np.array_equal(J0[ok], J1[ok]) and np.array_equal(h0[ok], h1[ok]) and np.array_equal(g0[ok], g1[ok])
        assert np.array_equal(out[0][6][ok], out[mode][6][ok])


@pytest.mark.parametrize("backend", BACKENDS)
def test_muller_ltrip_tilewalk_vs_per_step_and_c_twin(backend):
    # muller_2022, LTRIP(net) cluster graph (801 clusters / 1158 sepsets; sepsets of one or two nodes, so the
    # tile-walk kernel sees the (1,2), (2,1), (1,1) and copy shapes), 10 iterations over both spanning trees:
    # tile-walk (automatic) == per-step launches bit for bit, and both agree with the C twin of the oracle
    import bench
    from oracle.cport import COracle
    lib = get_lib(backend)
    w = bench.C3L()
    d = w.d
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)
    B = 5
    params, tips = w.inputs(B, 0)
    ref = COracle.from_plan_dict(d).run_batch(params, tips, root_belief=d["root_cluster"], want_fe=True, **w.cpu_kw)
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
    assert out[-1][2] * 10 < out[0][2]
    assert np.array_equal(out[-1][0], out[0][0]) and np.array_equal(out[-1][1], out[0][1])
    assert np.array_equal(out[-1][0], ref["iscal"])
    assert np.max(np.abs(out[-1][1] / ref["fe"] - 1)) <= 1e-10


@pytest.mark.parametrize("backend", BACKENDS)
def test_integratebelief_with_covariance(backend):
    # conditional moments of every belief after calibration: mean, inv(J), norm (the inputs of
    # calibrate_exact_cliquetree!, src/calibration.jl:462-463; test/test_exactBM.jl:26-52 checks them
    # against PhylogeneticEM through the same two calls)
    lib = get_lib(backend)
    p = 2
    R = np.array([[2.0, 0.5], [0.5, 1.0]])
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    rng = np.random.default_rng(11)
    B = 3
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p), np.diag([np.inf, np.inf]))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([R], np.zeros(p), np.diag([np.inf, np.inf])), data)
    assert bt.calibrate(case.sched)[0].all()
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
    for c in cgbs:
        OBP.calibrate(c, case.sched)
    for j in range(1, len(case.b) + 1):
        if bt.dimension(j) == 0:
            continue
        mu, cov, norm = bt.integratebelief_cov(j)
        mu2, norm2 = bt.integratebelief(j)
        assert np.array_equal(mu, mu2) and np.array_equal(norm, norm2)
        for e in range(B):
            ob = cgbs[e].belief[j - 1]
            assert relerr(cov[e], np.linalg.inv(ob.J)) <= TOL
            assert relerr(mu[e], np.linalg.solve(ob.J, ob.h)) <= TOL
            assert np.allclose(cov[e], cov[e].T, rtol=0, atol=0)


# ------------------------------------------------------------------ exact REML driver (calibrate_exact_cliquetree!)
@pytest.mark.parametrize("backend", BACKENDS)
def test_calibrate_exact_cliquetree_goldens(backend):
    # test/test_exactBM.jl:168-226: level-1 network, 4 taxa; univariate y and bivariate (x, y)
    lib = get_lib(backend)
    netstr = "(((A:4.0,((B1:1.0,B2:1.0)i6:0.6)#H5:1.1::0.9)i4:0.5,(#H5:2.0::0.1,C:0.1)i2:1.0)i1:3.0);"
    taxa = ["A", "B1", "B2", "C"]
    x = np.array([10.0, 10.0, 2.0, 0.0])
    y = np.array([1.0, 0.9, 1.0, -1.0])

    def plans(p):
        tbl = np.zeros((4, p))
        m_imp = M.MvFullBrownianMotion(np.eye(p), np.zeros(p), np.diag([np.inf] * p))
        m_fix = M.MvFullBrownianMotion(np.eye(p), np.zeros(p))
        c1 = Case(netstr, "cliquetree", tbl, taxa, m_imp, lib, schedule="spanningtree")
        c2 = Case(netstr, "cliquetree", tbl, taxa, m_fix, lib, schedule="spanningtree")
        return c1, c2

    c1, c2 = plans(1)
    data = np.stack([y[:, None], (2 * y + 1)[:, None]])  # second data set: affine image of the first
    s2, mu, ll = pgbp_b200.calibrate_exact_cliquetree(c1.plan, c2.plan, c1.sched[0], c2.sched[0], data)
    assert abs(ll[0] / -5.250084678427689 - 1) <= TOL
    assert abs(mu[0, 0] / -0.260008715071627 - 1) <= TOL
    assert abs(s2[0, 0, 0] / 0.4714735834478194 - 1) <= TOL
    # equivariance: y -> 2y + 1 gives mu -> 2 mu + 1, sigma2 -> 4 sigma2, loglik -> loglik - n log 2 (n = 4 tips)
    assert abs(mu[1, 0] - (2 * mu[0, 0] + 1)) <= TOL and abs(s2[1, 0, 0] / (4 * s2[0, 0, 0]) - 1) <= TOL
    assert abs((ll[1] - ll[0]) / (-4 * np.log(2)) - 1) <= TOL
    c1, c2 = plans(2)
    s2, mu, ll = pgbp_b200.calibrate_exact_cliquetree(c1.plan, c2.plan, c1.sched[0], c2.sched[0], np.stack([x, y], axis=1)[None])
    assert np.allclose(mu[0], [2.791001688545128, -0.260008715071627], rtol=TOL)
    assert np.allclose(s2[0], [[17.93326111121198, 1.6089749098736517], [1.6089749098736517, 0.4714735834478195]], rtol=TOL)
    assert np.isfinite(ll[0])


@pytest.mark.parametrize("backend", BACKENDS)
def test_calibrate_optimize_cliquetree_goldens(backend):
    # ML fit of UnivariateBrownianMotion by L-BFGS with a batched finite-difference stencil per iteration:
    # test/test_calibration.jl:242-244 (4-taxon level-1 network) and test/test_optimization.jl:16-18 (mateescu)
    lib = get_lib(backend)
    netstr = "(((A:4.0,((B1:1.0,B2:1.0)i6:0.6)#H5:1.1::0.9)i4:0.5,(#H5:2.0::0.1,C:0.1)i2:1.0)i1:3.0);"
    taxa = ["A", "B1", "B2", "C"]
    y = np.array([1.0, 0.9, 1.0, -1.0])[:, None]
    c = Case(netstr, "cliquetree", y, taxa, M.UnivariateBrownianMotion(1.0, -2.0), lib, schedule="spanningtree")
    theta, ll, res = pgbp_b200.calibrate_optimize_cliquetree(c.plan, c.sched[0], y, start=(1.0, -2.0))
    assert abs(ll / -5.174720533524127 - 1) <= TOL
    assert abs(theta[1] / -0.26000871507162693 - 1) <= 1e-5 and abs(theta[0] / 0.35360518758586457 - 1) <= 1e-5
    yd = np.array([1.0, -1.0])[:, None]  # tips d, g
    c = Case(GOLD["mateescu"], "cliquetree", yd, ["d", "g"], M.UnivariateBrownianMotion(1.0, 0.0), lib, schedule="spanningtree")
    theta, ll, res = pgbp_b200.calibrate_optimize_cliquetree(c.plan, c.sched[0], yd, start=(1.0, 0.0), maxiter=60)
    assert abs(ll / -3.2763180687070053 - 1) <= TOL
    assert abs(theta[0] / 0.5932930079336234 - 1) <= 1e-4 and abs(theta[1] / -0.07534357691418593 - 1) <= 1e-4


@pytest.mark.parametrize("backend", BACKENDS)
def test_calibrate_optimize_clustergraph_goldens(backend):
    # calibrate_optimize_clustergraph! on a loopy Bethe cluster graph (regularizebeliefs_bycluster!, calibrate!
    # with auto = true, objective = free energy): test/test_calibration.jl:188-204 (compared there with RxInfer,
    # rtol 1e-4) and test/test_optimization.jl:38-46 (mateescu_2010, against the exact ML fit)
    lib = get_lib(backend)
    y = np.array([11.275034507978296, 10.032494469945764, 11.49586603350308, 11.004447427824012])[:, None]
    c = Case(GOLD["netstr_unnamed"], "bethe", y, ["A", "B", "C", "D"], M.UnivariateBrownianMotion(1.0, 0.0), lib)
    theta, fe, res = pgbp_b200.calibrate_optimize_clustergraph(c.plan, c.sched, y, start=(1.0, 0.0), maxiter=100)
    assert abs(fe / -3.4312133894974126 - 1) <= 1e-4
    assert abs(theta[1] / 10.931640613828181 - 1) <= 1e-4 and abs(theta[0] / 0.15239159696122745 - 1) <= 1e-4
    yd = np.array([1.0, -1.0])[:, None]
    c = Case(GOLD["mateescu"], "bethe", yd, ["d", "g"], M.UnivariateBrownianMotion(1.0, 0.0), lib)
    theta, fe, res = pgbp_b200.calibrate_optimize_clustergraph(c.plan, c.sched, yd, start=(1.0, 0.0), maxiter=100)
    assert abs(theta[1] / -0.07534357691418593 - 1) <= 2e-5 and abs(theta[0] / 0.5932930079336234 - 1) <= 2e-6
    assert abs(fe / -3.2763180687070053 - 1) <= 3e-2


# ------------------------------------------------------------------ shared-precision batches (trait replicates under one theta)
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("p,method", [(3, "cliquetree"), (1, "bethe"), (6, "cliquetree")])
def test_shared_precision_batch_is_bit_identical(backend, p, method):
    # groups of elements that share a parameter vector store / update their J once (pgbp_batch_create_shared);
    # every result must equal the ordinary batch (each element its own J) bit for bit: beliefs, residuals,
    # flags, log-likelihood, factored energy, KL divergences, with and without regularisation
    lib = get_lib(backend)
    rng = np.random.default_rng(500 + p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    ntheta, nd = 3, 8
    B = ntheta * nd
    Rs = []
    for _ in range(ntheta):
        A = rng.normal(size=(p, p))
        Rs.append(A @ A.T / p + 0.1 * np.eye(p))
    mu = rng.normal(size=p)
    data = rng.normal(size=(nd, 7, p))
    model = M.MvFullBrownianMotion(Rs[0], mu)
    kw = dict(order_hint=GOLD["lazaridis_cluster_labels"]) if method == "cliquetree" else {}
    case = Case(GOLD["lazaridis"], method, data[0], taxa, model, lib, **kw)
    params = np.stack([pgbp_b200.bm_params([R], mu) for R in Rs])
    out = {}
    for name, group in (("own", 0), ("shared", nd)):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=group)
        bt.assignfactors(params, data, pairing="product")  # element e = (theta e // nd, data set e % nd)
        rec = {"assigned": [bt.get_belief(j) for j in range(1, len(case.b) + 1)]}
        if method == "bethe":
            bt.regularizebeliefs_bycluster()
        succ, iscal = bt.calibrate(case.sched, 2, update_residualkldiv=True)
        rec.update(succ=succ, iscal=iscal, st=bt.status(),
                   beliefs=[bt.get_belief(j) for j in range(1, len(case.b) + 1)],
                   res=[bt.get_residual(case.nclusters + 1 + j, case.plan.sepset_clusters[j][s] + 1)
                        for j in range(case.plan.nsepsets) for s in (0, 1)],
                   ll=bt.integratebelief(case.sched[0][2][0])[1], fe=bt.factored_energy(),
                   cov=bt.integratebelief_cov(case.sched[0][2][0]))
        out[name] = rec
    a, b_ = out["own"], out["shared"]
    assert a["succ"].all() and np.array_equal(a["succ"], b_["succ"]) and np.array_equal(a["iscal"], b_["iscal"])
    assert np.array_equal(a["st"], b_["st"])
    for key in ("assigned", "beliefs"):
        for x, y in zip(a[key], b_[key]):
            for u, v in zip(x, y):
                assert np.array_equal(u, v), key
    for x, y in zip(a["res"], b_["res"]):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)
    assert np.array_equal(a["ll"], b_["ll"]) and np.array_equal(a["fe"], b_["fe"])
    for u, v in zip(a["cov"], b_["cov"]):
        assert np.array_equal(u, v)
    # API restrictions of the shared mode
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=nd)
    with pytest.raises(pgbp_b200.PgbpError):
        bt.assignfactors(params[:1].repeat(B, axis=0), data[:1].repeat(B, axis=0))  # zip with B parameter sets
    bt.assignfactors(params, data, pairing="product")
    with pytest.raises(pgbp_b200.PgbpError):
        bt.calibrate(case.sched, 2, auto=True)
    with pytest.raises(pgbp_b200.PgbpError):
        pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=5)


@pytest.mark.parametrize("backend", BACKENDS)
def test_assignfactors_ou_device_vs_oracle_and_golden(backend):
    # UnivariateOrnsteinUhlenbeck factors assigned on the device (generic linear-Gaussian factor, one trait):
    # every cluster against the oracle's assignfactors!, log-likelihood against test/test_evomodels.jl:120
    # (-42.31401134496844 for OU(2, 3, -2, 0.0, 0.4) on trait y), for random, fixed and improper roots
    lib = get_lib(backend)
    tbl = TBL[:, [1]]
    recs = [(2.0, 3.0, -2.0, 0.0, 0.4), (0.7, 0.4, 1.5, -0.3, 0.0), (1.3, 1.1, 0.2, 0.0, math.inf)]
    for k, (s2, al, th, mu, v) in enumerate(recs):
        model = M.UnivariateOrnsteinUhlenbeck(s2, al, th, mu, None if v == 0.0 else v)
        case = Case(GOLD["netstr_named"], "cliquetree", tbl, TAXA, model, lib, schedule="spanningtree")
        B = 3
        data = np.stack([tbl, tbl + 0.25, 2 * tbl])
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
        bt.assignfactors_ou([[s2, al, th, mu, v]], data)
        assert (bt.status() == 0).all()
        cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
        check_all_beliefs(case, bt, cgbs)
        spt = case.sched[0]
        assert bt.propagate_1traversal_postorder(spt).all()
        ll = bt.integratebelief(spt[2][0])[1]
        for e in range(B):
            assert OBP.propagate_1traversal_postorder(cgbs[e], *spt)
            assert abs(ll[e] / OBP.integratebelief_cgb(cgbs[e], spt[2][0])[1] - 1) <= TOL
        if k == 0:
            assert abs(ll[0] / -42.31401134496844 - 1) <= TOL


# ------------------------------------------------------------------ regressions (round-1 advisor findings)
@pytest.mark.parametrize("backend", BACKENDS)
def test_two_batches_one_plan_propagate_over_edge_outside_every_tree(backend):
    # The plan is immutable and shared: propagate_belief! over an edge that is in no spanning tree must not grow
    # the plan's index tables (a second batch created earlier would index past its device copy).  Plan without
    # any tree: every message descriptor comes from pgbp_propagate.
    lib = get_lib(backend)
    p = 2
    R = np.array([[1.0, 0.3], [0.3, 2.0]])
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    rng = np.random.default_rng(19)
    data = rng.normal(size=(2, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    fam = None
    plan = pgbp_b200.ClusterGraphPlan(case.plan.nclusters, case.plan.belief_dim, case.plan.sepset_clusters,
                                      case.plan.upind, [], p, fam, lib)
    A = pgbp_b200.BatchedClusterGraphBelief(plan, 2)
    Bb = pgbp_b200.BatchedClusterGraphBelief(plan, 2)
    cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(2)]
    case.upload(A, cgbs)
    case.upload(Bb, cgbs)
    spt = case.sched[0]
    par, chi = spt[2][-1], spt[3][-1]  # a real edge of the clique tree (1-based cluster indices)
    sep = None
    for j, (a, b_) in enumerate(plan.sepset_clusters):
        if {a + 1, b_ + 1} == {par, chi}:
            sep = plan.nclusters + 1 + j
    assert sep is not None
    A.propagate_belief(par, sep, chi)
    Bb.propagate_belief(par, sep, chi)   # crashed (out-of-bounds table read) before the plan became immutable
    A.propagate_belief(chi, sep, par)
    Bb.propagate_belief(chi, sep, par)
    for j in (par, chi, sep):
        for u, v in zip(A.get_belief(j), Bb.get_belief(j)):
            assert np.array_equal(u, v)
    for e in range(2):
        bl = cgbs[e].belief
        assert OBP.propagate_belief(bl[par - 1], bl[sep - 1], bl[chi - 1]) is None
        assert OBP.propagate_belief(bl[chi - 1], bl[sep - 1], bl[par - 1]) is None
    J, h, g = A.get_belief(par)
    for e in range(2):
        ob = cgbs[e].belief[par - 1]
        assert relerr(J[e], ob.J) <= TOL and relerr(h[e], ob.h) <= TOL and abs(g[e] - ob.g) <= TOL * max(1, abs(ob.g))


@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_failed_leader_does_not_stall_its_group(backend):
    # A per-element failure on a group LEADER (here: NaN tip data, status set by factor assignment) must not stop
    # the J updates of its group: J depends on the shared parameters only.  The other elements of the group must
    # get exactly the results of an ordinary batch.
    lib = get_lib(backend)
    p = 3
    rng = np.random.default_rng(77)
    A_ = rng.normal(size=(p, p))
    R = A_ @ A_.T / p + 0.1 * np.eye(p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    B, gs = 8, 4
    data = rng.normal(size=(B, 7, p))
    data[4, 2, 1] = np.nan  # element 4 leads the second group
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    params = pgbp_b200.bm_params([R], np.zeros(p))
    out = {}
    for name, group in (("own", 0), ("shared", gs)):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=group)
        bt.assignfactors(params, data)
        succ, iscal = bt.calibrate(case.sched, 1)
        out[name] = (succ, bt.status(), bt.integratebelief(case.sched[0][2][0])[1], bt.factored_energy(),
                     [bt.get_belief(j) for j in (1, 5, case.nclusters + 2)])
    keep = np.arange(B) != 4
    so, ss = out["own"], out["shared"]
    assert not so[0][4] and not ss[0][4] and so[0][keep].all() and ss[0][keep].all()
    assert (ss[1][keep] == 0).all() and ss[1][4] != 0
    assert np.array_equal(so[2][keep], ss[2][keep]) and np.array_equal(so[3][keep], ss[3][keep])
    for x, y in zip(so[4], ss[4]):
        # J of the failed leader's group is still the group's J (read through the leader's column)
        assert np.array_equal(x[0][5:], y[0][5:]) and np.array_equal(x[1][keep], y[1][keep])


# ------------------------------------------------------------------ PhylogeneticEM moments (test/test_exactBM.jl:20-52)
@pytest.mark.parametrize("backend", BACKENDS)
def test_phylogeneticem_conditional_moments_goldens(backend):
    # test/test_exactBM.jl:3-52: 5-taxon tree, trait y, UnivariateBrownianMotion(1, 0, 1e10) ("infinite" root variance to
    # match PhylogeneticEM), clique tree; PhylogeneticEM's log-likelihood, conditional expectations, variances and
    # parent-child covariances at EVERY belief, through integratebelief! + inv(J) = pgbp_integrate_cov.
    # The reference holds 7 digits (atol 1e-6); the same moments from a dense conditional-Gaussian computation on the
    # tree covariance (independent of belief propagation) agree to 2e-6 absolute (the dense route cancels against the 1e10 root
    # variance), and the oracle's calibrated beliefs pin the device moments to 1e-10.
    lib = get_lib(backend)
    netstr = "((A:1.5,B:1.5):1,(C:1,(D:0.5, E:0.5):0.5):1.5);"
    taxa = ["A", "B", "C", "D", "E"]
    y = np.array([1.0, 0.9, 1.0, -1.0, -0.9])
    v0 = 1e10
    model = M.UnivariateBrownianMotion(1.0, 0.0, v0)
    case = Case(netstr, "cliquetree", y[:, None], taxa, model, lib, schedule="spanningtree")
    B = 2
    data = np.stack([y[:, None], y[:, None]])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([1.0], [0.0], [v0]), data)
    assert bt.calibrate(case.sched)[0].all()
    llscore = -18.83505
    condexp = np.array([1, 0.9, 1, -1, -0.9, 0.4436893, 0.7330097, 0.009708738, -0.6300971])[[5, 7, 8, 4, 3, 2, 6, 1, 0]]
    condvar = np.array([0, 0, 0, 0, 0, 0.9174757, 0.5970874, 0.3786408, 0.2087379])[[5, 7, 8, 4, 3, 2, 6, 1, 0]]
    condcov = np.array([0, 0, 0, 0, 0, np.nan, 0.3932039, 0.2038835, 0.1262136])[[5, 7, 8, 4, 3, 2, 6, 1, 0]]
    # dense check values: joint Gaussian of all 9 nodes (preorder), conditioned on the 5 tips
    nodes = case.net.vec_node
    n = len(nodes)
    idx = case.net.preorder_index()
    V = np.zeros((n, n))
    V[0, 0] = v0
    for i in range(1, n):
        (e,) = nodes[i].parent_edges()
        q = idx[id(e.parent)] - 1
        V[i, :i] = V[q, :i]
        V[:i, i] = V[i, :i]
        V[i, i] = V[q, q] + e.length
    tip = np.array([i for i in range(n) if nodes[i].leaf])
    internal = np.array([i for i in range(n) if not nodes[i].leaf])
    ytip = np.array([y[taxa.index(nodes[i].name)] for i in tip])
    Vtt = V[np.ix_(tip, tip)]
    K = np.linalg.solve(Vtt, V[np.ix_(tip, internal)]).T
    cmean = np.zeros(n)
    cmean[tip] = ytip
    cmean[internal] = K @ ytip
    ccov = np.zeros((n, n))
    ccov[np.ix_(internal, internal)] = V[np.ix_(internal, internal)] - K @ V[np.ix_(tip, internal)]
    sign, logdet = np.linalg.slogdet(Vtt)
    ll_dense = -0.5 * (len(tip) * np.log(2 * np.pi) + logdet + ytip @ np.linalg.solve(Vtt, ytip))
    assert abs(ll_dense - llscore) <= 1e-5 and np.allclose(cmean, condexp, atol=1e-6) and np.allclose(np.diag(ccov), condvar, atol=1e-6)
    cgb = case.oracle_cgb(tbl=y[:, None])
    OBP.calibrate(cgb, case.sched)
    for j in range(1, len(case.b) + 1):
        if bt.dimension(j) == 0:
            continue
        mu, cov, norm = bt.integratebelief_cov(j)
        ob = cgb.belief[j - 1]
        for e in range(B):
            assert relerr(cov[e], np.linalg.inv(ob.J)) <= TOL and relerr(mu[e], np.linalg.solve(ob.J, ob.h)) <= TOL
            assert abs(norm[e] / OBP.integratebelief_cgb(cgb, j)[1] - 1) <= TOL
        lab = [int(v) for v in case.b[j - 1].nodelabel if not case.n2x[int(v) - 1]]  # in-scope nodes (preorder, 1-based)
        assert len(lab) == bt.dimension(j)
        for e in range(B):
            # the reference's assertions (atol 1e-6 against PhylogeneticEM)
            assert abs(mu[e, -1] - condexp[lab[-1] - 1]) <= 1e-6
            assert abs(norm[e] - llscore) <= 1e-5
            assert abs(cov[e, -1, -1] - condvar[lab[-1] - 1]) <= 1e-6
            if len(lab) == 2 and not np.isnan(condcov[lab[0] - 1]):
                assert abs(cov[e, 0, 1] - condcov[lab[0] - 1]) <= 1e-6
            # and against the dense conditional Gaussian (which loses ~7 digits to the 1e10 root variance)
            assert abs(norm[e] / ll_dense - 1) <= 1e-7
            li = [v - 1 for v in lab]
            assert np.max(np.abs(mu[e] - cmean[li])) <= 2e-6 and np.max(np.abs(cov[e] - ccov[np.ix_(li, li)])) <= 2e-6


@pytest.mark.parametrize("backend", BACKENDS)
def test_root_status_update_fixed_vs_random_plans(backend):
    # test/test_exactBM.jl:95-165: beliefs allocated for a random root and re-allocated for a fixed root
    # (init_beliefs_allocate_atroot!, host side: the plan of the other root status is simply built from the re-allocated
    # beliefs) must give the same factors as beliefs allocated for that root status from the start.  Here: the plan built
    # for one root status and a plan built from scratch for the other, both assigned on the device, against the oracle's
    # assignfactors! for the respective model -- every belief, for the univariate and the diagonal bivariate model.
    lib = get_lib(backend)
    netstr = "((A:1.5,B:1.5):1,(C:1,(D:0.5, E:0.5):0.5):1.5);"
    taxa = ["A", "B", "C", "D", "E"]
    tbl = np.array([[10, 1.0], [10, 0.9], [3, 1.0], [0, -1.0], [1, -0.9]])
    for cols, rates, mu, v in (([1], [1.0], [0.0], [0.9]), ([0, 1], [1.0, 1.0], [0.0, 0.0], [1.2, 3.0])):
        p = len(cols)
        data = tbl[:, cols]
        mk = (lambda vv: M.UnivariateBrownianMotion(rates[0], mu[0], vv[0] if vv else None)) if p == 1 else \
             (lambda vv: M.MvDiagBrownianMotion(rates, mu, vv if vv else None))
        m_rand, m_fix = mk(v), mk(None)
        c_rand = Case(netstr, "cliquetree", data, taxa, m_rand, lib, schedule="spanningtree")
        c_fix = Case(netstr, "cliquetree", data, taxa, m_fix, lib, schedule="spanningtree")
        # scopes: the fixed root leaves the scope of every belief, nothing else changes
        for b1, b2 in zip(c_rand.b, c_fix.b):
            assert list(b1.nodelabel) == list(b2.nodelabel)
        assert sum(c_rand.plan.belief_dim) == sum(c_fix.plan.belief_dim) + p * sum(1 for b_ in c_rand.b if 1 in list(b_.nodelabel))
        for case, model, vv in ((c_rand, m_rand, v), (c_fix, m_fix, None)):
            bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 2)
            par = pgbp_b200.bm_params([np.diag(rates) if p > 1 else rates[0]], mu, (np.diag(vv) if p > 1 else vv) if vv else None)
            bt.assignfactors(par, np.stack([data, data]))
            cgb = case.oracle_cgb(tbl=data, model=model)
            check_all_beliefs(case, bt, [cgb, cgb])
            assert bt.calibrate(case.sched)[0].all()
            OBP.calibrate(cgb, case.sched)
            check_all_beliefs(case, bt, [cgb, cgb])
        # same data, same tree: the random-root likelihood tends to the fixed-root one as the root variance vanishes


# ------------------------------------------------------------------ regularisation on a real network (docs/src/man/regularization.md:150-201)
@pytest.mark.parametrize("backend", BACKENDS)
def test_lipson_bethe_regularisation_behaviour(backend):
    # lipson_2020b (44 nodes, 11 hybrids), Bethe cluster graph, UnivariateBrownianMotion(1, 0): one iteration of
    # calibrate! WITHOUT regularisation meets ill-defined messages (the documented ones: belief H5I5I16 integrating
    # [2, 3] in the postorder pass) -> per-element status naming that message; after regularizebeliefs_bynodesubtree!
    # or regularizebeliefs_onschedule! there are none, and every belief equals the oracle's.
    lib = get_lib(backend)
    taxa, x = GOLD["lipson_taxa"], np.array(GOLD["lipson_x"])[:, None]
    model = M.UnivariateBrownianMotion(1.0, 0.0)
    case = Case(GOLD["lipson"], "bethe", x, taxa, model, lib)
    assert len(case.net.vec_node) == 44 and sum(n.leaf for n in case.net.vec_node) == 12
    B = 3
    data = np.stack([x, x + 0.1, 0.5 * x])
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    bt.assignfactors(pgbp_b200.bm_params([1.0], [0.0]), data)
    succ, iscal = bt.calibrate(case.sched, 1)
    assert not succ.any() and not iscal.any()
    st = bt.status()
    # the oracle's first failing message of the first postorder traversal
    cgb = case.oracle_cgb(tbl=x)
    spt = case.sched[0]
    n = len(spt[0])
    first = None
    for r, i in enumerate(range(n - 1, -1, -1)):
        sep = cgb.belief[cgb.sepsetindex(spt[0][i], spt[1][i]) - 1]
        flag = OBP.propagate_belief(cgb.belief[spt[2][i] - 1], sep, cgb.belief[spt[3][i] - 1])
        if flag is not None:
            first = (r, spt[1][i], flag.info)
            break
    assert first is not None
    lab, integ = GOLD["lipson_bethe_failing_beliefs"][0]
    assert first[1] == lab  # docs: "belief H5I5I16, integrating [2, 3]"
    sender = cgb.belief[case.cg.labels.index(lab)]
    assert sender.dimension() == 3 and integ == [2, 3]
    for e in range(B):  # status = (position of the message in the reference's sequential order, LAPACK info)
        assert (st[e] >> 8) - 1 == first[0] and (st[e] & 0xff) == first[2]
    # with regularisation: no ill-defined message, beliefs equal to the oracle's
    for kind in ("bynodesubtree", "onschedule"):
        bt.clear_status()
        bt.init_beliefs_reset_fromfactors()
        bt.init_messagecalibrationflags_reset()
        cgbs = [case.oracle_cgb(tbl=data[e]) for e in range(B)]
        if kind == "bynodesubtree":
            bt.regularizebeliefs_bynodesubtree(OBP.bynodesubtree_program(cgbs[0], case.cg))
            for c in cgbs:
                OBP.regularizebeliefs_bynodesubtree(c, case.cg)
        else:
            bt.regularizebeliefs_onschedule()
            for c in cgbs:
                OBP.regularizebeliefs_onschedule(c, case.cg)
        succ, _ = bt.calibrate(case.sched, 1)
        assert succ.all() and (bt.status() == 0).all(), kind
        for c in cgbs:
            assert OBP.calibrate(c, case.sched, 1)[0]
        check_all_beliefs(case, bt, cgbs, tol=TOL)


# ------------------------------------------------------------------ shared-precision batches: factored layout, wider coverage
@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_layout_memory_and_rows(backend):
    # J rows once per group: device memory of a shared batch ~ (h, g rows) x B + one group batch, far below an ordinary
    # batch; pgbp_batch_belief_rows names the compact rows of the device view
    lib = get_lib(backend)
    p = 6
    rng = np.random.default_rng(8)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    B = 512
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    own = pgbp_b200.BatchedClusterGraphBelief(case.plan, B)
    sh = pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=B)
    assert sh.device_bytes() < own.device_bytes() / 4
    dims = case.plan.belief_dim
    row = 0
    for j in range(1, len(dims) + 1):
        hrow, grow = sh.belief_rows(j)
        assert (hrow, grow) == (row, row + dims[j - 1])
        row += dims[j - 1] + 1
    base, ld, nrows = sh.device_view()
    assert nrows == row and ld >= B
    h_own, g_own = own.belief_rows(3)
    assert (h_own, g_own) == case.plan.belief_slot(2)[1:]


@pytest.mark.parametrize("B,jmsg_wide", [(6, None), (200, None), (6, "1")])
@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_large_shapes_and_single_messages(backend, B, jmsg_wide, monkeypatch):
    # p = 16 on a small synthetic level-1 network (sender dimensions 16 / 32 / 48: the compile-time I = 16 / 32 element
    # kernels and the warp-cooperative group kernel on 48 x 48 matrices), and the single-message paths
    # (propagate_belief!, regularizebeliefs_onschedule!) -- everything bit-identical to an ordinary batch.
    # B = 200: on the GPU the element pass of the I = 16 / 32 messages goes through the bulk-copy kernel (k_hmsg_bulk:
    # blocks of 128 elements, the second one ragged).  PGBP_JMSG_WIDE=1: the group pass takes the one-warp kernel for
    # every launch (what it does on its own from 1,024 (message, group) pairs per launch upwards)
    if jmsg_wide:
        monkeypatch.setenv("PGBP_JMSG_WIDE", jmsg_wide)
    lib = get_lib(backend)
    import bench
    w = bench.C5(ntips=60, nretic=6)
    d = w.d
    params, tips = w.inputs(B, 0)
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)
    out = {}
    for name, group in (("own", 0), ("shared", B)):
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, shared_precision_group=group)
        bt.assignfactors(params, tips)
        succ, iscal = bt.calibrate(None, 1)
        root = d["root_cluster"] + 1
        rec = dict(succ=succ, iscal=iscal, ll=bt.integratebelief(root)[1], fe=bt.factored_energy(),
                   b=[bt.get_belief(j) for j in (1, 2, root, plan.nclusters + 1, plan.nclusters + plan.nsepsets)])
        # one more message by hand over the first tree edge, both directions, then a regularised restart
        par, chi = d["trees"][0][0][0] + 1, d["trees"][0][1][0] + 1
        sep = next(plan.nclusters + 1 + j for j, ab in enumerate(plan.sepset_clusters) if set(ab) == {par - 1, chi - 1})
        bt.propagate_belief(par, sep, chi)
        bt.propagate_belief(chi, sep, par)
        rec["p"] = [bt.get_belief(j) for j in (par, chi, sep)]
        bt.init_beliefs_reset_fromfactors()
        bt.init_messagecalibrationflags_reset()
        bt.regularizebeliefs_onschedule()
        succ2, _ = bt.calibrate(None, 1)
        rec["r"] = [bt.get_belief(j) for j in (1, root, plan.nclusters + 1)] + [(succ2,)]
        rec["st"] = bt.status()
        out[name] = rec
    a, s_ = out["own"], out["shared"]
    assert a["succ"].all() and np.array_equal(a["succ"], s_["succ"]) and np.array_equal(a["iscal"], s_["iscal"])
    assert np.array_equal(a["ll"], s_["ll"]) and np.array_equal(a["fe"], s_["fe"]) and np.array_equal(a["st"], s_["st"])
    for key in ("b", "p", "r"):
        for x, y in zip(a[key], s_[key]):
            for u, v in zip(x, y):
                assert np.array_equal(u, v), key


@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_missing_data_and_ou(backend):
    # trait-level scopes (missing data: scoped K1 path, all-zero shortcut of marginalize) and the OU model on
    # shared-precision batches: equal to ordinary batches bit for bit
    lib = get_lib(backend)
    m = M.MvDiagBrownianMotion([2, 1], [3, -3], [0.1, 10])
    case = Case(GOLD["netstr_named"], "cliquetree", TBL, TAXA, m, lib, schedule="spanningtree")
    B = 4
    rng = np.random.default_rng(21)
    data = np.repeat(TBL[None], B, axis=0)
    data[1:] += 0.2 * rng.normal(size=(B - 1,) + TBL.shape)
    par = pgbp_b200.bm_params([np.diag([2.0, 1.0])], [3, -3], np.diag([0.1, 10.0]))
    res = {}
    for name, group in (("own", 0), ("shared", 2)):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=group)
        bt.assignfactors(par, data)
        succ, _ = bt.calibrate(case.sched)
        res[name] = (succ, bt.status(), bt.integratebelief(case.sched[0][2][0])[1],
                     [bt.get_belief(j) for j in range(1, len(case.b) + 1)])
    assert res["own"][0].all() and np.array_equal(res["own"][1], res["shared"][1])
    assert np.array_equal(res["own"][2], res["shared"][2])
    assert abs(res["own"][2][0] / -21.347496753649892 - 1) <= TOL  # test/test_evomodels.jl:190
    for x, y in zip(res["own"][3], res["shared"][3]):
        for u, v in zip(x, y):
            assert np.array_equal(u, v, equal_nan=True)
    # OU, one trait
    tbl = TBL[:, [1]]
    model = M.UnivariateOrnsteinUhlenbeck(2.0, 3.0, -2.0, 0.0, 0.4)
    case = Case(GOLD["netstr_named"], "cliquetree", tbl, TAXA, model, lib, schedule="spanningtree")
    data = np.stack([tbl, tbl + 0.25, 2 * tbl, tbl - 1])
    res = {}
    for name, group in (("own", 0), ("shared", 4)):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, 4, shared_precision_group=group)
        bt.assignfactors_ou([[2.0, 3.0, -2.0, 0.0, 0.4]], data)
        assert bt.calibrate(case.sched)[0].all()
        res[name] = (bt.integratebelief(case.sched[0][2][0])[1], bt.factored_energy())
    assert np.array_equal(res["own"][0], res["shared"][0]) and np.array_equal(res["own"][1], res["shared"][1])
    assert abs(res["own"][0][0] / -42.31401134496844 - 1) <= TOL


@pytest.mark.parametrize("backend", BACKENDS)
def test_calibrate_optimize_cliquetree_mvfull_closed_form(backend):
    # calibrate_optimize_cliquetree! for MvFullBrownianMotion through the log-Cholesky transform of
    # params_optimize / params_original (src/evomodels/homogeneousbrownianmotion.jl:130-159, src/calibration.jl:182-234)
    # on a network with complete bivariate data, against the closed-form ML fit of a matrix-normal model:
    # mu = (1'V^-1 1)^-1 1'V^-1 Y, R = (Y - 1 mu)'V^-1 (Y - 1 mu) / n with V the network's tip covariance at unit rate
    from oracle import densemvn
    lib = get_lib(backend)
    netstr = "(((A:4.0,((B1:1.0,B2:1.0)i6:0.6)#H5:1.1::0.9)i4:0.5,(#H5:2.0::0.1,C:0.1)i2:1.0)i1:3.0);"
    taxa = ["A", "B1", "B2", "C"]
    Y = np.array([[10.0, 1.0], [10.0, 0.9], [2.0, 1.0], [0.0, -1.0]])
    n, p = Y.shape
    c = Case(netstr, "cliquetree", Y, taxa, M.MvFullBrownianMotion(np.eye(p), np.zeros(p)), lib, schedule="spanningtree")
    C1 = densemvn.network_covariance(c.net, lambda e: np.eye(1), 1)
    idx = {v.name: i for i, v in enumerate(c.net.vec_node)}
    sel = [idx[t] for t in taxa]
    V = C1[np.ix_(sel, sel)]
    Vi = np.linalg.inv(V)
    one = np.ones(n)
    mu_ml = (one @ Vi @ Y) / (one @ Vi @ one)
    E = Y - mu_ml
    R_ml = E.T @ Vi @ E / n
    ll_ml = -0.5 * (n * p * np.log(2 * np.pi) + p * np.linalg.slogdet(V)[1] + n * np.linalg.slogdet(R_ml)[1] + n * p)
    theta, ll, res = pgbp_b200.calibrate_optimize_cliquetree(c.plan, c.sched[0], Y, model="MvFullBrownianMotion",
                                                              start=(np.array([[2.0, 0.3], [0.3, 1.0]]), np.array([3.0, 0.0])),
                                                              maxiter=200)
    R_hat = theta[:p * p].reshape(p, p)
    mu_hat = theta[p * p:p * p + p]
    assert abs(ll / ll_ml - 1) <= 1e-8
    assert np.allclose(R_hat, R_ml, rtol=2e-3, atol=0) and np.allclose(mu_hat, mu_ml, rtol=2e-3, atol=1e-4)
    # the transform round-trips
    from pgbp_b200.drivers import _bm_transforms
    to_opt, to_orig = _bm_transforms("MvFullBrownianMotion", p, None)
    th = to_opt(R_ml, mu_ml)
    assert th.size == p * (p + 1) // 2 + p
    back = to_orig(th)
    assert np.allclose(back[:p * p].reshape(p, p), R_ml, rtol=1e-13) and np.allclose(back[p * p:p * p + p], mu_ml, rtol=1e-13)


@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_wide_network_two_groups(backend):
    # a network with > 1024 clusters and only two groups: on the GPU the group batch's K1 goes through the row-parallel
    # launcher (one thread per cluster) and the group pass sees levels of hundreds of messages; product pairing with
    # two parameter vectors x four data sets; heterogeneous colours through the K1 block cache
    lib = get_lib(backend)
    import bench
    w = bench.C4(ntips=700, nretic=70, p=3)
    d = w.d
    assert d["nclusters"] > 1024
    nth, nd = 2, 4
    B = nth * nd
    params, _ = w.inputs(nth, 0)
    tips = w.synth.simulate_tips(d, lambda v, k: np.eye(3), nd, 99)
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)
    root = d["root_cluster"] + 1
    out = {}
    for name, group in (("own", 0), ("shared", nd)):
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, shared_precision_group=group)
        bt.assignfactors(params, tips, ncolors=w.ncolors, pairing="product")
        succ, iscal = bt.calibrate(None, 1)
        out[name] = (succ, iscal, bt.status(), bt.integratebelief(root)[1], bt.factored_energy(),
                     [bt.get_belief(j) for j in (1, 17, root, plan.nclusters + 5)])
    a, s_ = out["own"], out["shared"]
    assert a[0].all() and all(np.array_equal(a[k], s_[k]) for k in range(5))
    for x, y in zip(a[5], s_[5]):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)


@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_groups_of_128_elements(backend):
    # two parameter vectors x 128 data sets, p = 8 (integrated dimensions 8 / 16): every block of 128 elements lies
    # inside one group, which is the condition for the bulk-copy element pass on the GPU (k_hmsg_bulk), here with
    # residual tracking and two iterations (second traversal reads non-zero sepsets); bit-identical to own-J batches
    lib = get_lib(backend)
    import bench
    w = bench.C4(ntips=40, nretic=4, p=8)
    d = w.d
    nth, nd = 2, 128
    B = nth * nd
    params, _ = w.inputs(nth, 0)
    tips = w.synth.simulate_tips(d, lambda v, k: np.eye(8), nd, 7)
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)
    root = d["root_cluster"] + 1
    out = {}
    for name, group in (("own", 0), ("shared", nd)):
        bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, shared_precision_group=group)
        bt.assignfactors(params, tips, ncolors=w.ncolors, pairing="product")
        succ, iscal = bt.calibrate(None, 2, update_residualkldiv=True)
        out[name] = (succ, iscal, bt.status(), bt.integratebelief(root)[1], bt.factored_energy(),
                     [bt.get_belief(j) for j in (1, 5, root, plan.nclusters + 3)]
                     + [bt.get_residual(plan.nclusters + 1 + j, plan.sepset_clusters[j][s] + 1)
                        for j in (0, 2) for s in (0, 1)])
    a, s_ = out["own"], out["shared"]
    assert a[0].all() and all(np.array_equal(a[k], s_[k]) for k in range(5))
    for x, y in zip(a[5], s_[5]):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)


@pytest.mark.parametrize("backend", BACKENDS)
def test_shared_precision_large_batch_one_group(backend):
    # 8,320 replicates under one parameter vector on the lazaridis clique tree (p = 3): 65 blocks of 128 elements, the
    # first size at which the walk kernels step aside for per-step launches (and, with PGBP_SHARED_WALK=1, the element
    # walk runs behind the group walk's event); bit-identical to own-J batches
    lib = get_lib(backend)
    rng = np.random.default_rng(77)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    p, B = 3, 8320
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    mu = rng.normal(size=p)
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, mu)
    case = Case(GOLD["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=GOLD["lazaridis_cluster_labels"])
    params = pgbp_b200.bm_params([R], mu)[None, :]
    out = {}
    for name, group in (("own", 0), ("shared", B)):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, B, shared_precision_group=group)
        bt.assignfactors(params, data, pairing="product")
        succ, iscal = bt.calibrate(case.sched, 2)
        out[name] = (succ, iscal, bt.status(), bt.integratebelief(case.sched[0][2][0])[1], bt.factored_energy(),
                     [bt.get_belief(j) for j in range(1, len(case.b) + 1)])
    a, s_ = out["own"], out["shared"]
    assert a[0].all() and all(np.array_equal(a[k], s_[k]) for k in range(5))
    for x, y in zip(a[5], s_[5]):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)
