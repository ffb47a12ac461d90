"""Pins the CPU oracle against the reference's own known answers.

Every number below is a golden recorded in the reference's test suite or
jldoctests (file:line cited per test, relative to the reference checkout).
Reference tests use `≈` (rtol sqrt(eps)); here rtol 1e-9 unless the reference
itself states a looser one.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import beliefs as B
from oracle import bp
from oracle import clustergraph as CG
from oracle import densemvn
from oracle import models as M
from oracle.network import preprocessnet, readnewick

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
NAN = math.nan
NETSTR_NAMED = GOLD["netstr_named"]
NETSTR_UNNAMED = GOLD["netstr_unnamed"]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def setup(netstr, method, tbl, taxa, model, **kw):
    net = readnewick(netstr)
    cg = CG.clustergraph(net, method, **kw)
    b, (n2c, n2f, n2x, n2d, c2n) = B.allocatebeliefs(tbl, taxa, net.vec_node, cg, model)
    B.assignfactors(b, model, tbl, taxa, net.vec_node, n2c, n2f, n2x)
    cgb = bp.ClusterGraphBelief(b, n2c, n2f, n2x, c2n)
    return net, cg, b, cgb


# ---------------------------------------------------------------- network / graph
def test_preorder_and_names():
    # test/test_evomodels.jl:156
    net = readnewick(NETSTR_NAMED)
    preprocessnet(net)
    assert [n.name for n in net.vec_node] == GOLD["preorder_named"]


def test_minfill_order_and_cliquetree():
    # test/test_clustergraph.jl:8-13, 120-134
    net = readnewick(GOLD["netstr_cg"])
    preprocessnet(net)
    g = CG.moralize(net)
    assert g.nv() == len(net.nodes) and g.ne() == len(net.edges) + 1
    assert CG.triangulate_minfill(g) == GOLD["minfill_order_cg"]
    assert g.ne() == 13
    ct = CG.cliquetree(g)
    assert ct.ne() == 8 and CG.is_tree(ct)
    assert sorted(ct.sepset(*l) for l in ct.edge_labels()) == GOLD["cliquetree_sepsets_cg"]
    assert all(t[1] for t in CG.check_runningintersection(ct, net))
    assert CG.isfamilypreserving([v[1] for v in ct.vdata.values()], net)[0]
    net = readnewick(GOLD["mateescu"])
    ct = CG.clustergraph(net, "cliquetree")
    assert CG.is_tree(ct) and ct.vdata["H3DH1B"][1] == [5, 4, 3, 2]


def test_bethe_ltrip_jgs_structure():
    # test/test_clustergraph.jl:43-118
    net = readnewick(GOLD["netstr_cg"])
    cg = CG.clustergraph(net, "bethe")
    ntaxa = sum(n.leaf for n in net.nodes)
    nhyb = sum(n.hybrid for n in net.nodes)
    assert cg.nv() == (len(net.nodes) - 1) + (len(net.nodes) - ntaxa)
    nint = sum(not e.hybrid for e in net.edges) - ntaxa
    assert cg.ne() == ntaxa + 2 * nint + 3 * nhyb
    assert CG.is_connected(cg)
    assert all(t[1] for t in CG.check_runningintersection(cg, net))
    assert sorted(v[1] for v in cg.vdata.values()) == sorted(
        [[1], [3], [4], [6], [8], [9], [2, 1], [3, 1], [4, 3], [5, 4], [6, 4], [7, 6], [8, 3],
         [9, 8, 6], [10, 9], [11, 8]])
    clusters = [[11, 8], [10, 9], [7, 6], [5, 4], [2, 1], [9, 8, 6], [8, 3], [6, 4], [4, 3], [3, 1]]
    cg = CG.clustergraph(net, "ltrip", clusters=clusters)
    assert sorted(clusters) == sorted(v[1] for v in cg.vdata.values())
    assert CG.is_connected(cg)
    assert all(t[1] for t in CG.check_runningintersection(cg, net))
    cg = CG.clustergraph(net, "ltrip")
    assert all(t[1] for t in CG.check_runningintersection(cg, net))
    with pytest.raises(ValueError):
        CG.clustergraph(net, "ltrip", clusters=[[11, 8], [10, 9], [7, 6], [5, 4], [2, 1], [9, 8],
                                                [8, 3], [6, 4], [4, 3], [3, 1]])
    net = readnewick(GOLD["mateescu"])
    cg = CG.clustergraph(net, "jgs", maxclustersize=3)
    assert all(t[1] for t in CG.check_runningintersection(cg, net))
    assert not CG.is_tree(cg)
    assert sorted(v[1] for v in cg.vdata.values()) == GOLD["jgs3_clusters_mateescu"]
    assert sorted(cg.sepset(*l) for l in cg.edge_labels()) == GOLD["jgs3_sepsets_mateescu"]
    with pytest.raises(ValueError):
        CG.clustergraph(net, "jgs", maxclustersize=2)


def test_spanningtrees_cover_all_edges():
    # test/test_clustergraph.jl:132-148
    net = readnewick(GOLD["netstr_cg"])
    cg = CG.clustergraph(net, "bethe")
    sched = CG.spanningtrees_clusterlist(cg, net.vec_node)
    covered = set()
    for spt in sched:
        assert len(spt[0]) == cg.nv() - 1
        sg, _ = CG.induced_subgraph_edges(cg, [(min(a, b), max(a, b)) for a, b in zip(spt[2], spt[3])])
        assert set(sg.labels) == set(cg.labels) and CG.is_tree(sg)
        covered |= {cg.arrange(a, b) for a, b in zip(spt[0], spt[1])}
    assert covered == set(cg.edata)


def test_lazaridis_docs_listing():
    # docs/src/man/getting_started.md:107-125 (labels), :160-163 (node labels),
    # :184-189 (J, g of belief 1), :245-261 (schedule), :283-291 (norm, fe)
    net = readnewick(GOLD["lazaridis"])
    ct = CG.clustergraph(net, "cliquetree", order_hint=GOLD["lazaridis_cluster_labels"])
    assert ct.labels == GOLD["lazaridis_cluster_labels"]
    taxa = net.tiplabels()
    assert taxa == ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    tbl = np.array(GOLD["lazaridis_x"]).reshape(-1, 1)
    m = M.UnivariateBrownianMotion(1, 0)
    b, (n2c, n2f, n2x, n2d, c2n) = B.allocatebeliefs(tbl, taxa, net.vec_node, ct, m)
    assert len(b) == 33
    assert b[0].nodelabel == [17, 16, 10]
    B.assignfactors(b, m, tbl, taxa, net.vec_node, n2c, n2f, n2x)
    np.testing.assert_allclose(b[0].J, np.array(GOLD["lazaridis_b1_J"]), rtol=1e-13)
    assert rel(b[0].g, GOLD["lazaridis_b1_g"]) < 1e-13
    cgb = bp.ClusterGraphBelief(b, n2c, n2f, n2x, c2n)
    assert cgb.nclusters == 17 and cgb.nsepsets() == 16
    sched = CG.spanningtrees_clusterlist(ct, net.vec_node)
    assert len(sched) == 1
    assert [ct.labels.index(l) + 1 for l in sched[0][1]] == GOLD["lazaridis_sched_child"]
    assert sched[0][2] == GOLD["lazaridis_sched_parent"]
    assert bp.calibrate(cgb, sched) == (True, False)  # 1 pass: residuals not yet small
    _, norm = bp.integratebelief_inplace(b[0])
    assert rel(norm, GOLD["lazaridis_norm"]) < 1e-12
    fe = bp.factored_energy(cgb)[2]
    assert rel(fe, GOLD["lazaridis_fe"]) < 1e-12
    # ML fit end point as a fixed-theta known answer (:322-331)
    m = M.UnivariateBrownianMotion(0.31812948798414614, 1.1525789703018783)
    B.assignfactors(b, m, tbl, taxa, net.vec_node, n2c, n2f, n2x)
    assert bp.propagate_1traversal_postorder(cgb, *sched[0])
    _, ll = bp.integratebelief_cgb(cgb, sched[0][2][0])
    assert rel(ll, -8.656529929205751) < 1e-9


# ---------------------------------------------------------------- canonical form
def test_canonicalform_beliefs_and_six_messages():
    # test/test_canonicalform.jl:44-63 (node labels / scope), :65-109
    tbl = np.array([[10, 1.0], [10, 0.9], [NAN, 1.0], [0, -1.0]])
    taxa = ["A", "B1", "B2", "C"]
    labels = GOLD["canonicalform_beliefnodelabels"]
    hint = labels[:7]
    net = readnewick(NETSTR_NAMED)
    ct = CG.clustergraph(net, "cliquetree", order_hint=hint)
    b, _ = B.allocatebeliefs(tbl, taxa, net.vec_node, ct, M.UnivariateBrownianMotion(1, 0, 1))
    assert [be.nodelabel for be in b] == labels
    assert b[4].inscope.all() and b[4].inscope.shape == (2, 3)
    assert list(B.scopeindex_nodes([6], b[0])) == [0, 1]
    assert B.scopeindex_nodes([5], b[0]).size > 0
    with pytest.raises(ValueError):
        B.scopeindex_nodes([2], b[0])
    b, _ = B.allocatebeliefs(tbl, taxa, net.vec_node, ct, M.UnivariateBrownianMotion(1, 0, 0))
    assert (b[4].inscope == np.array([[1, 1, 0], [1, 1, 0]], dtype=bool)).all()

    tbl_y = tbl[:, 1:2]
    m = M.UnivariateBrownianMotion(2, 3, 0)
    b, (n2c, n2f, n2x, n2d, c2n) = B.allocatebeliefs(tbl_y, taxa, net.vec_node, ct, m)
    B.assignfactors(b, m, tbl_y, taxa, net.vec_node, n2c, n2f, n2x)
    e = {k.number: k for k in net.edges}
    mJ = 1 / 2
    np.testing.assert_allclose(b[0].J, mJ / e[4].length * np.array([[1, -1], [-1, 1]]))
    assert (b[0].h == 0).all()
    assert rel(b[0].g, -math.log(2 * math.pi * e[4].length * 2) / 2) < 1e-14
    bpv = mJ / e[3].length
    np.testing.assert_allclose(b[1].J, [[bpv]])
    np.testing.assert_allclose(b[1].h, [bpv * 1.0])
    assert rel(b[1].g, -(math.log(2 * math.pi / bpv) + bpv * 1.0 ** 2) / 2) < 1e-14
    bpv = mJ / e[2].length
    np.testing.assert_allclose(b[2].h, [bpv * 0.9])
    bpv = mJ / (e[7].gamma ** 2 * e[7].length + e[5].gamma ** 2 * e[5].length)
    np.testing.assert_allclose(b[3].J, bpv * np.array([[1, -.9, -.1], [-.9, .81, .09], [-.1, .09, .01]]))
    assert rel(b[3].g, -math.log(2 * math.pi / bpv) / 2) < 1e-14
    bp2 = mJ / np.array([e[6].length, e[9].length])
    np.testing.assert_allclose(b[4].J, np.diag(bp2))
    np.testing.assert_allclose(b[4].h, bp2 * 3)
    assert rel(b[4].g, -np.sum(np.log(2 * math.pi / bp2) + bp2 * 9) / 2) < 1e-14
    for to, s, fr in [(1, 8, 2), (1, 9, 3), (4, 10, 1), (4, 12, 6), (4, 13, 7), (5, 11, 4)]:
        ss = b[s - 1]
        assert bp.propagate_belief(b[to - 1], ss, b[fr - 1], B.MessageResidual(ss.J, ss.h)) is None
    _, ll = bp.integratebelief_inplace(b[4])
    assert ll == pytest.approx(GOLD["canonicalform_loglik"], rel=1e-14)


# ---------------------------------------------------------------- evolutionary models
@pytest.mark.parametrize("case", GOLD["evomodels"], ids=lambda c: c["id"])
def test_evomodels_postorder_loglik(case):
    # test/test_evomodels.jl:52-264 (one postorder + integratebelief! at the root cluster)
    tbl = np.array([[10, 1.0], [10, 0.9], [NAN, 1.0], [0, -1.0]])
    taxa = ["A", "B1", "B2", "C"]
    cols = {"y": [1], "x": [0], "xy": [0, 1]}[case["traits"]]
    inf = math.inf
    args = eval(case["args"], {"inf": inf, "np": np})
    model = getattr(M, case["model"])(*args)
    net, ct, b, cgb = setup(NETSTR_NAMED, "cliquetree", tbl[:, cols], taxa, model)
    spt = CG.spanningtree_clusterlist(ct, prenodes=net.vec_node)
    assert bp.propagate_1traversal_postorder(cgb, *spt)
    _, ll = bp.integratebelief_cgb(cgb, spt[2][0])
    assert rel(ll, case["loglik"]) < 1e-9


def test_evomodels_against_dense_mvn():
    tbl = np.array([[10, 1.0], [10, 0.9], [NAN, 1.0], [0, -1.0]])
    taxa = ["A", "B1", "B2", "C"]
    net = readnewick(NETSTR_NAMED)
    preprocessnet(net)
    R = np.array([[2.0, 0.5], [0.5, 1.0]])
    V = np.array([[0.1, 0.01], [0.01, 0.2]])
    assert rel(densemvn.loglik_bm(net, tbl, taxa, lambda e: R, [3.0, -3.0]), -24.312323855394055) < 1e-12
    assert rel(densemvn.loglik_bm(net, tbl, taxa, lambda e: R, [3.0, -3.0], rootvar=V), -23.16482738327936) < 1e-12
    assert rel(densemvn.loglik_bm(net, tbl, taxa, lambda e: R, [3.0, -3.0], improper=True), -16.9626044836951) < 1e-9
    assert rel(densemvn.loglik_bm(net, tbl[:, 1:], taxa, lambda e: np.array([[2.0]]), [3.0]), -10.732857817537196) < 1e-13


# ---------------------------------------------------------------- calibration
def test_calibration_cliquetree_improper_root_and_regularisation():
    # test/test_calibration.jl:36-78
    tbl_y = np.array([[1.0], [.9], [1], [-1]])
    taxa = ["A", "B1", "B2", "C"]
    m = M.UnivariateBrownianMotion(0.471474, 0, math.inf)
    net, ct, b, ctb = setup(NETSTR_NAMED, "cliquetree", tbl_y, taxa, m)
    spt = CG.spanningtree_clusterlist(ct, prenodes=net.vec_node)
    bp.calibrate(ctb, [spt])
    ll = -4.877930583154144
    for i in range(1, len(b) + 1):
        assert bp.integratebelief_cgb(ctb, i)[1] == pytest.approx(ll, rel=1e-7)
    assert bp.factored_energy(ctb)[2] == pytest.approx(ll, rel=1e-7)
    root_ind = next(i for i, be in enumerate(b) if 1 in be.nodelabel)
    assert bp.integratebelief_inplace(b[root_ind])[0][-1] == pytest.approx(-0.26000871507162693, rel=1e-5)
    assert np.linalg.inv(b[root_ind].J)[-1, -1] == pytest.approx(0.33501871740664146, rel=1e-5)
    for reg in (bp.regularizebeliefs_bynodesubtree, bp.regularizebeliefs_bycluster):
        bp.init_beliefs_reset_fromfactors(ctb)
        reg(ctb, ct)
        bp.calibrate(ctb, [spt])
        assert bp.integratebelief_cgb(ctb, 1)[1] == pytest.approx(ll, rel=1e-7)


def test_calibration_bethe_onschedule():
    # test/test_calibration.jl:79-105: converges at iteration 5, schedule tree 1
    tbl_y = np.array([[-1.81358], [0.468158], [0.658486], [0.643821]])
    taxa = ["A", "B", "C", "D"]
    m = M.UnivariateBrownianMotion(0.0861249, 0)
    net, cg, b, cgb = setup(NETSTR_UNNAMED, "bethe", tbl_y, taxa, m)
    bp.regularizebeliefs_onschedule(cgb, cg)
    sched = CG.spanningtrees_clusterlist(cg, net.vec_node)
    assert all(bp.calibrate(cgb, sched, 20, auto=True))
    assert bp.calibrate.last_info == (5, 1)
    ind = cgb.clusterindex("I3")
    assert bp.integratebelief_inplace(b[ind - 1])[0][-1] == pytest.approx(0.21511454631828986, rel=1e-5)


def test_calibration_missing_data_tree():
    # test/test_calibration.jl:107-129
    tbl = np.array([[1, NAN], [1, NAN], [1, NAN], [1, 1.0]])
    m = M.MvDiagBrownianMotion([1, 1], [0, 0])
    net, ct, b, ctb = setup("(((A:1.0, B:1.0)E:1.0, C:2.0)F:1.0, D:3.0)G;", "cliquetree", tbl,
                            ["A", "B", "C", "D"], m)
    spt = CG.spanningtree_clusterlist(ct, prenodes=net.vec_node)
    bp.calibrate(ctb, [spt])
    for i in range(1, len(b) + 1):
        assert bp.integratebelief_cgb(ctb, i)[1] == pytest.approx(-7.578343735986344, rel=1e-9)


def test_calibration_joingraph_bynodesubtree():
    # test/test_calibration.jl:131-185
    netstr = GOLD["netstr_level3"]
    tbl = np.array([[2.11, 30.0], [2.15, NAN]])
    taxa = ["A", "B"]
    m = M.MvFullBrownianMotion([[1, 0.5], [0.5, 1]], [0, 0], [[math.inf, 0], [0, math.inf]])
    net, cg, b, cgb = setup(netstr, "jgs", tbl, taxa, m, maxclustersize=3)
    bp.regularizebeliefs_bynodesubtree(cgb, cg)
    sch = []
    for n in net.vec_node:
        st = CG.nodesubtree_clusterlist(cg, n.name)
        if st[0]:
            sch.append(st)
    assert all(bp.calibrate(cgb, sch, 10, auto=True))
    i6 = cgb.clusterindex("I1I2I3")
    mu, nrm = bp.integratebelief_inplace(b[i6 - 1])
    assert nrm == pytest.approx(-1.390595772423, rel=1e-7)
    np.testing.assert_allclose(mu, [2.121105154896223, 30.005552577448075, 2.1360649504455984,
                                    30.013032475222563, 2.128585052670943, 30.00929252633547], rtol=1e-7)
    i2 = cgb.clusterindex("H1H2I1")
    mu, nrm = bp.integratebelief_inplace(b[i2 - 1])
    assert nrm == pytest.approx(-1.390595772423, rel=1e-7)
    np.testing.assert_allclose(mu, [2.125583120364, 30.007791560181964, 2.129918967774073,
                                    30.009959483886966, 2.121105154896214, 30.00555257744811], rtol=1e-7)
    m = M.MvFullBrownianMotion([[1, 0.5], [0.5, 1]], [2.128585052670943, 30.00929252633547])
    b, (n2c, n2f, n2x, n2d, c2n) = B.allocatebeliefs(tbl, taxa, net.vec_node, cg, m)
    B.assignfactors(b, m, tbl, taxa, net.vec_node, n2c, n2f, n2x)
    cgb = bp.ClusterGraphBelief(b, n2c, n2f, n2x, c2n)
    bp.regularizebeliefs_bynodesubtree(cgb, cg)
    assert all(bp.calibrate(cgb, sch, 10, auto=True))
    mu, nrm = bp.integratebelief_inplace(b[cgb.clusterindex("I1I2I3") - 1])
    assert nrm == pytest.approx(-3.3498677834866997, rel=1e-7)
    np.testing.assert_allclose(mu, [2.121105154896223, 30.005552577448075, 2.1360649504455984,
                                    30.013032475222563], rtol=1e-7)


def test_residual_kldiv():
    # test/test_calibration.jl:13-33
    res = B.MessageResidual(np.zeros((2, 2)), np.zeros(2))
    res.dJ[:] = np.ones((2, 2)) / 3
    res.dh[:] = np.array([-2, 4]) / 3
    sep = B.CanonicalBelief([1, 2], 1, np.ones((1, 2), bool), B.SEPSET, ("A", "B"))
    sep.J[:] = np.eye(2)
    sep.h[:] = [0, 1]
    bp.residual_kldiv_update(res, sep)
    assert res.kldiv == pytest.approx(1.215973, rel=1e-6)


def test_fixed_theta_optimiser_endpoints():
    # test/test_optimization.jl:16-18 (mateescu); test/test_calibration.jl:242-244, 279-281
    net = readnewick(GOLD["mateescu"])
    preprocessnet(net)
    taxa = ["d", "g"]
    tbl = np.array([[1.0], [-1.0]])
    m = M.UnivariateBrownianMotion(0.5932930079336234, -0.07534357691418593)
    net, ct, b, ctb = setup(GOLD["mateescu"], "cliquetree", tbl, taxa, m)
    spt = CG.spanningtree_clusterlist(ct, prenodes=net.vec_node)
    assert bp.propagate_1traversal_postorder(ctb, *spt)
    assert bp.integratebelief_cgb(ctb, spt[2][0])[1] == pytest.approx(-3.2763180687070053, rel=1e-9)
    tbl = np.array([[10, 1.0], [10, 0.9], [NAN, 1.0], [0, -1.0]])
    taxa = ["A", "B1", "B2", "C"]
    m = M.UnivariateBrownianMotion(0.35360518758586457, -0.26000871507162693)
    net, ct, b, ctb = setup(NETSTR_NAMED, "cliquetree", tbl[:, 1:], taxa, m)
    spt = CG.spanningtree_clusterlist(ct, prenodes=net.vec_node)
    bp.propagate_1traversal_postorder(ctb, *spt)
    assert bp.integratebelief_cgb(ctb, spt[2][0])[1] == pytest.approx(-5.174720533524127, rel=1e-9)
    m = M.MvDiagBrownianMotion([11.257682945973125, 0.35360518758586457],
                               [3.500266520382341, -0.26000871507162693])
    net, ct, b, ctb = setup(NETSTR_NAMED, "cliquetree", tbl, taxa, m)
    bp.propagate_1traversal_postorder(ctb, *spt)
    assert bp.integratebelief_cgb(ctb, spt[2][0])[1] == pytest.approx(-14.39029465611705, rel=1e-9)


def test_muller_2022_structure_goldens():
    # docs/src/man/clustergraphs.md:40-41 (network), :52-56 (clique tree), :101-117 (Bethe):
    # pins the Newick reader / preorder / moralisation / min-fill / Bethe construction at 801 nodes
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "muller_2022.json")))
    net = readnewick(g["newick"])
    assert len(net.nodes) == g["nnodes"] and len(net.edges) == g["nedges"]
    assert sum(n.leaf for n in net.nodes) == g["ntips"] and sum(n.hybrid for n in net.nodes) == g["nhybrids"]
    fg = CG.clustergraph(net, "bethe")
    assert len(fg.labels) == g["bethe"]["nclusters"] and fg.ne() == g["bethe"]["nsepsets"]
    sizes = [len(fg.vdata[l][1]) for l in fg.labels]
    assert max(sizes) == g["bethe"]["max_cluster_size"] and abs(np.mean(sizes) - g["bethe"]["mean_cluster_size"]) < 1e-6
    ct = CG.clustergraph(net, "cliquetree")
    assert len(ct.labels) == g["cliquetree"]["nclusters"] and ct.ne() == g["cliquetree"]["nsepsets"]
    sizes = [len(ct.vdata[l][1]) for l in ct.labels]
    assert min(sizes) == 2 and max(sizes) == 54  # docs/src/man/clustergraphs.md:73-89
    sched = CG.spanningtrees_clusterlist(fg, net.vec_node)
    covered = {frozenset((a, b)) for spt in sched for a, b in zip(spt[2], spt[3])}
    assert len(covered) == fg.ne()  # every edge of the loopy graph is on some spanning tree
