"""Dumps the plan inputs of the benchmark workloads as JSON.

In production these arrays come from the Julia side (cluster graph, scopeindex
maps, spanning-tree schedules, node families: INTEGRATION.md).  Julia is not in
the build image, so they are generated ONCE here with the oracle's Julia-free
front-end and committed; bench.py's GPU arm only reads the JSON (it never
imports oracle/).  Re-run:  python workloads/make_plans.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import beliefs as OB  # noqa: E402
from oracle import clustergraph as CG  # noqa: E402
from oracle import models as M  # noqa: E402
from oracle.network import readnewick  # noqa: E402
import pgbp_b200  # noqa: E402
from harness import prenodes_info  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")))


def dump(name, netstr, method, p, taxa, order_hint=None, **kw):
    net = readnewick(netstr)
    cg = CG.clustergraph(net, method, order_hint=order_hint, **kw) if order_hint else CG.clustergraph(net, method, **kw)
    taxa = taxa or net.tiplabels()
    model = M.MvFullBrownianMotion(np.eye(p), np.zeros(p))  # fixed root, no missing data: scopes only
    tbl = np.zeros((len(taxa), p))
    b, (n2c, n2f, n2x, n2d, c2n) = OB.allocatebeliefs(tbl, taxa, net.vec_node, cg, model)
    nc = len(cg.labels)
    sched = CG.spanningtrees_clusterlist(cg, net.vec_node)
    fam = pgbp_b200.families_table(prenodes_info(net), n2c, n2f, n2x, b, p, True, taxa)
    lab2idx = {l: i for i, l in enumerate(cg.labels)}
    sc, up = [], []
    for s in b[nc:]:
        a, b_ = lab2idx[s.metadata[0]], lab2idx[s.metadata[1]]
        sc.append([a, b_])
        up.append([[int(x) for x in OB.scopeindex(s, b[a])], [int(x) for x in OB.scopeindex(s, b[b_])]])
    out = dict(name=name, network=netstr if len(netstr) < 2000 else "(see tests/golden)", method=method, ntraits=p, taxa=taxa, nclusters=nc,
               cluster_labels=cg.labels, belief_dim=[x.dimension() for x in b], sepset_clusters=sc, upind=up,
               trees=[[[j - 1 for j in t[2]], [j - 1 for j in t[3]]] for t in sched], families=fam,
               root_cluster=sched[0][2][0] - 1,
               # node families for simulating traits down the network: per node (preorder), parents as
               # (parent index, length, gamma)
               simulate=[[[q[0] - 1, q[1], q[2]] for q in info["parents"]] for info in prenodes_info(net)],
               tip_nodes=[[i for i, n in enumerate(net.vec_node) if n.name == t][0] for t in taxa])
    path = os.path.join(ROOT, "workloads", name + ".json")
    json.dump(out, open(path, "w"))
    print(path, "clusters", nc, "sepsets", len(b) - nc, "trees", len(sched))


if __name__ == "__main__":
    dump("lazaridis_cliquetree_p3", GOLD["lazaridis"], "cliquetree", 3,
         ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"],
         order_hint=GOLD["lazaridis_cluster_labels"])
    # BASELINE configs[2]: muller_2022 (801 nodes, 40 tips, 361 hybrids), Bethe cluster graph
    # (1557 clusters / 1914 sepsets), univariate, spanningtrees_clusterlist schedule (2 trees)
    MULLER = json.load(open(os.path.join(ROOT, "tests", "golden", "muller_2022.json")))
    dump("muller_bethe_p1", MULLER["newick"], "bethe", 1, None)
    # same network, LTRIP(net) cluster graph (801 clusters / 1158 sepsets, docs/src/man/clustergraphs.md:131-132)
    dump("muller_ltrip_p1", MULLER["newick"], "ltrip", 1, None)
