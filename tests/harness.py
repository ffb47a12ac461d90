"""Test harness: builds product-side plans/batches from oracle-side objects.

The oracle provides the Julia-free front-end (network -> cluster graph ->
beliefs' scopes -> schedules); the product consumes that output exactly as it
would consume the reference's in production."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pgbp_b200  # noqa: E402
from oracle import beliefs as OB  # noqa: E402
from oracle import bp as OBP  # noqa: E402
from oracle import clustergraph as CG  # noqa: E402
from oracle.network import readnewick  # noqa: E402

PKG = os.path.join(ROOT, "phylogaussianbeliefprop.jl_b200")
EMUL = os.path.join(PKG, "lib", "libpgbp_emul.so")
CUDA = os.path.join(PKG, "lib", "libpgbp_b200.so")

BACKENDS = [pytest.param("emul", id="emul"), pytest.param("cuda", marks=pytest.mark.gpu, id="cuda")]
_libs = {}


def get_lib(backend):
    """emul: host-emulation build of the kernel bodies (CPU tests of the host
    logic only); cuda: the product."""
    if backend not in _libs:
        if backend == "emul":
            if not os.path.exists(EMUL):
                import importlib.util
                spec = importlib.util.spec_from_file_location("pgbp_build", os.path.join(PKG, "build.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                mod.build(emul=True)
            _libs[backend] = pgbp_b200.Library(EMUL)
        else:
            import torch
            assert torch.cuda.is_available(), "gpu test without a CUDA device"
            _libs[backend] = pgbp_b200.Library(CUDA)
    return _libs[backend]


def prenodes_info(net):
    idx = net.preorder_index()
    out = []
    for n in net.vec_node:
        out.append(dict(name=n.name, leaf=n.leaf,
                        parents=[(idx[id(e.parent)], e.length, e.gamma, e.number) for e in n.parent_edges()]))
    return out


class Case:
    """One network + cluster graph + model scope, oracle side and product side."""

    def __init__(self, netstr, method, tbl, taxa, model, lib, schedule="spanningtrees", with_families=True,
                 edge_color=None, cg=None, **kw):
        self.net = readnewick(netstr) if isinstance(netstr, str) else netstr
        self.cg = cg if cg is not None else CG.clustergraph(self.net, method, **kw)
        self.tbl = np.asarray(tbl, dtype=float)
        self.taxa = list(taxa)
        self.model = model
        b, (n2c, n2f, n2x, n2d, c2n) = OB.allocatebeliefs(self.tbl, self.taxa, self.net.vec_node, self.cg, model)
        self.b, self.n2c, self.n2f, self.n2x, self.c2n = b, n2c, n2f, n2x, c2n
        self.nclusters = len(self.cg.labels)
        if schedule == "spanningtrees":
            self.sched = CG.spanningtrees_clusterlist(self.cg, self.net.vec_node)
        elif schedule == "spanningtree":
            self.sched = [CG.spanningtree_clusterlist(self.cg, prenodes=self.net.vec_node)]
        else:
            self.sched = schedule(self)
        fam = None
        if with_families:
            tm = np.isnan(np.atleast_2d(self.tbl).reshape(len(self.taxa), -1)) if np.isnan(self.tbl).any() else None
            fam = pgbp_b200.families_table(prenodes_info(self.net), n2c, n2f, n2x, b, model.ntraits,
                                           model.isrootfixed(), self.taxa, edge_color, tip_missing=tm)
        self.plan = pgbp_b200.ClusterGraphPlan.from_beliefs(b, self.nclusters, self.cg.labels, self.sched, fam, lib)

    def oracle_cgb(self, tbl=None, model=None):
        """Fresh oracle ClusterGraphBelief with factors assigned."""
        tbl = self.tbl if tbl is None else tbl
        model = self.model if model is None else model
        b = [x.copy() for x in self.b]
        OB.assignfactors(b, model, tbl, self.taxa, self.net.vec_node, self.n2c, self.n2f, self.n2x)
        return OBP.ClusterGraphBelief(b, self.n2c, self.n2f, self.n2x, self.c2n)

    def upload(self, batch, cgbs):
        """Upload the beliefs of a list of oracle ClusterGraphBeliefs (one per element)."""
        for j in range(len(self.b)):
            m = self.b[j].dimension()
            J = np.stack([c.belief[j].J for c in cgbs]) if m else np.zeros((len(cgbs), 0, 0))
            h = np.stack([c.belief[j].h for c in cgbs]) if m else np.zeros((len(cgbs), 0))
            g = np.array([c.belief[j].g for c in cgbs])
            batch.set_belief(j + 1, J, h, g)


def synth_oracle_objects(net_tab, plan):
    """Oracle-side Network + cluster graph (MetaGraph) + schedule of a network / clique tree made by
    workloads/synth.py, so that the whole reference-following oracle pipeline (allocatebeliefs,
    assignfactors!, calibrate!, ...) can run on it.  Node names are n<preorder index>."""
    from oracle.network import Edge, Network, Node
    n = net_tab["nnodes"]
    net = Network()
    nodes = [Node(name=f"n{v}", leaf=bool(net_tab["leaf"][v]), hybrid=len(net_tab["parents"][v]) > 1) for v in range(n)]
    for v in range(n):
        for (q, L, g, eno) in net_tab["parents"][v]:
            e = Edge(number=eno, length=L, gamma=g, hybrid=len(net_tab["parents"][v]) > 1, child=nodes[v], parent=nodes[q])
            net.edges.append(e)
    # edge lists: children first (in child order), then parent edges, like the Newick reader
    for e in net.edges:
        e.parent.edges.append(e)
    for e in net.edges:
        e.child.edges.append(e)
    net.nodes = list(nodes)
    net.root = nodes[0]
    net.vec_node = list(nodes)
    cg = CG.MetaGraph("cliquetree")
    labs = []
    for nodes_c in plan["cluster_nodes"]:
        lab = "".join(f"n{v}" for v in nodes_c)
        labs.append(lab)
        cg.add_vertex(lab, ([f"n{v}" for v in nodes_c], [v + 1 for v in nodes_c]))
    for (a, b), sn in zip(plan["sepset_clusters"], plan["sepset_nodes"]):
        cg.add_edge(labs[a], labs[b], [v + 1 for v in sn])
    tp, tc = plan["trees"][0]
    sched = [([labs[a] for a in tp], [labs[b] for b in tc], [a + 1 for a in tp], [b + 1 for b in tc])]
    taxa = [f"n{v}" for v in plan["tip_nodes"]]
    return net, cg, sched, taxa


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    d = np.where(np.isnan(d) & (np.isnan(a) == np.isnan(b)) & (np.isinf(a) == np.isinf(b)), 0.0, d)
    return float(np.max(d)) / scale
