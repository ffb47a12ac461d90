"""Cluster graphs and message schedules -- restatement of src/clustergraph.jl.

TEST INFRASTRUCTURE (see oracle/__init__.py).  In production this layer stays
in Julia; its OUTPUT FORMAT (cluster order, sepset order, 4-vector spanning
trees) is the input contract of the hot path, so
it is restated here to let the tests build real inputs without Julia.

Third-party behaviour restated (Graphs.jl 1.x / MetaGraphsNext 0.7, not
vendored under /root/reference; pinned by docs/src/man/getting_started.md:
245-261 -- the 16-edge lazaridis schedule -- and test/test_clustergraph.jl):
  * SimpleGraph adjacency lists are sorted; `edges(g)` iterates (src<dst)
    lexicographically; `edge_labels` follows it;
  * `kruskal_mst` sorts edge weights with a stable sort (ties keep edge order);
  * `dfs_parents` is an iterative DFS taking the first unseen neighbour;
  * `topological_sort` is `topological_sort_by_dfs` started from vertex 1,2,..
    descending into the LAST unvisited out-neighbour, finishing order reversed
    (the variant that reproduces the documented lazaridis schedule exactly);
  * `induced_subgraph(g, edgelist)` renumbers vertices in order of appearance;
  * `rem_vertex!` moves the last vertex into the freed slot.
NOT restated: the iteration order of Julia `Set`/`Dict` (hash dependent).  It
decides the ORDER of `maximal_cliques` (hence cluster indices of a clique
tree) and of minibuckets in join-graph structuring.  Here cliques come in a
canonical order unless `order_hint` pins the order documented by the reference
(used by the golden tests).  Log-likelihoods do not depend on that order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

from .network import Network, nodefamilies, preprocessnet


# --------------------------------------------------------------------------
# a tiny labelled simple graph with MetaGraphsNext-like semantics
# --------------------------------------------------------------------------
class MetaGraph:
    """Undirected labelled graph.  Vertex codes are 1-based positions in
    `labels`.  vdata[label] = (node_names, node_preorder_indices);
    edata[(l1,l2)] (labels arranged by code) = list of preorder indices."""

    def __init__(self, tag=""):
        self.tag = tag
        self.labels: List[str] = []
        self.vdata: Dict[str, object] = {}
        self.adj: List[set] = []  # adj[code-1] = set of neighbour codes
        self.edata: Dict[Tuple[str, str], object] = {}

    # vertices
    def nv(self):
        return len(self.labels)

    def ne(self):
        return len(self.edata)

    def code_for(self, lab):
        return self.labels.index(lab) + 1

    def label_for(self, code):
        return self.labels[code - 1]

    def has_vertex(self, lab):
        return lab in self.vdata

    def add_vertex(self, lab, data):
        if lab in self.vdata:
            return False
        self.labels.append(lab)
        self.vdata[lab] = data
        self.adj.append(set())
        return True

    def arrange(self, l1, l2):
        return (l1, l2) if self.code_for(l1) < self.code_for(l2) else (l2, l1)

    def has_edge(self, l1, l2):
        return self.arrange(l1, l2) in self.edata

    def add_edge(self, l1, l2, data):
        """MetaGraphsNext add_edge!: if the edge exists its data is replaced."""
        c1, c2 = self.code_for(l1), self.code_for(l2)
        if c1 == c2:
            return False
        self.adj[c1 - 1].add(c2)
        self.adj[c2 - 1].add(c1)
        self.edata[self.arrange(l1, l2)] = data
        return True

    def rem_edge(self, l1, l2):
        c1, c2 = self.code_for(l1), self.code_for(l2)
        self.adj[c1 - 1].discard(c2)
        self.adj[c2 - 1].discard(c1)
        self.edata.pop(self.arrange(l1, l2), None)

    def delete_vertex(self, lab):
        """rem_vertex! semantics: last vertex takes the freed code."""
        c = self.code_for(lab)
        for nb in list(self.adj[c - 1]):
            self.rem_edge(lab, self.label_for(nb))
        n = self.nv()
        if c != n:
            last = self.labels[n - 1]
            old = {k: v for k, v in self.edata.items() if last in k}
            for k in old:
                del self.edata[k]
            nbrs = self.adj[n - 1]
            for nb in nbrs:
                self.adj[nb - 1].discard(n)
                self.adj[nb - 1].add(c)
            self.adj[c - 1] = set(nbrs)
            self.labels[c - 1] = last
            for (a, b), v in old.items():
                other = b if a == last else a
                self.edata[self.arrange(last, other)] = v
        self.labels.pop()
        self.adj.pop()
        del self.vdata[lab]

    def neighbors(self, code):
        return sorted(self.adj[code - 1])

    def neighbor_labels(self, lab):
        return [self.label_for(c) for c in self.neighbors(self.code_for(lab))]

    def edges(self):
        """(src,dst) code pairs, src<dst, lexicographic -- Graphs.edges order."""
        out = []
        for s in range(1, self.nv() + 1):
            for d in sorted(self.adj[s - 1]):
                if d > s:
                    out.append((s, d))
        return out

    def edge_labels(self):
        return [(self.label_for(s), self.label_for(d)) for s, d in self.edges()]

    def sepset(self, l1, l2):
        return self.edata[self.arrange(l1, l2)]

    def copy(self):
        g = MetaGraph(self.tag)
        g.labels = list(self.labels)
        g.vdata = dict(self.vdata)
        g.adj = [set(a) for a in self.adj]
        g.edata = {k: (list(v) if isinstance(v, list) else v) for k, v in self.edata.items()}
        return g


def is_tree(g: MetaGraph):
    return g.ne() == g.nv() - 1 and is_connected(g)


def is_connected(g: MetaGraph):
    if g.nv() == 0:
        return True
    seen, st = {1}, [1]
    while st:
        v = st.pop()
        for u in g.adj[v - 1]:
            if u not in seen:
                seen.add(u)
                st.append(u)
    return len(seen) == g.nv()


# --------------------------------------------------------------------------
# moralize / triangulate  (src/clustergraph.jl:44-121)
# --------------------------------------------------------------------------
def moralize(net: Network) -> MetaGraph:
    """src/clustergraph.jl:44-77.  vdata = preorder index."""
    idx = net.preorder_index()
    g = MetaGraph("moralized")
    for code, n in enumerate(net.vec_node, start=1):
        g.add_vertex(n.name, code)
    for e in net.edges:
        g.add_edge(e.parent.name, e.child.name, "hybrid" if e.hybrid else "tree")
    for n in net.nodes:
        if not n.hybrid:
            continue
        pl = [p.name for p in n.parents()]
        for i1 in range(len(pl)):
            for i2 in range(i1 + 1, len(pl)):
                if not g.has_edge(pl[i1], pl[i2]):
                    g.add_edge(pl[i1], pl[i2], "moralized")
    return g


def triangulate_minfill(graph: MetaGraph) -> List[str]:
    """src/clustergraph.jl:87-121: greedy min-fill, ties -> largest preorder
    index; `graph` gains the fill edges; returns the elimination order."""
    # work on adjacency by label (vertex codes are irrelevant: scores are unique)
    adj = {lab: {graph.label_for(c) for c in graph.adj[i]} for i, lab in enumerate(graph.labels)}
    pre = dict(graph.vdata)
    ordering = []
    while len(adj) > 1:
        best, bestscore = None, None
        for v, nb in adj.items():
            nbl = list(nb)
            fill = 0
            for a in range(len(nbl)):
                na = adj[nbl[a]]
                for b in range(a + 1, len(nbl)):
                    if nbl[b] not in na:
                        fill += 1
            sc = (fill, -pre[v])
            if bestscore is None or sc < bestscore:
                best, bestscore = v, sc
        nbl = sorted(adj[best], key=lambda l: pre[l])
        for a in range(len(nbl)):
            for b in range(a + 1, len(nbl)):
                if nbl[b] not in adj[nbl[a]]:
                    adj[nbl[a]].add(nbl[b])
                    adj[nbl[b]].add(nbl[a])
                    graph.add_edge(nbl[a], nbl[b], "fill")
        ordering.append(best)
        for nb in adj[best]:
            adj[nb].discard(best)
        del adj[best]
    ordering.append(next(iter(adj)))
    return ordering


def maximal_cliques_chordal(graph: MetaGraph) -> List[List[int]]:
    """Maximal cliques of a chordal graph, as lists of vertex CODES.  Computed
    from a perfect elimination order (maximum cardinality search); canonical
    order = by largest preorder index in the clique, decreasing.  (The
    reference calls Graphs.maximal_cliques, src/clustergraph.jl:760, whose
    output order depends on Julia Set iteration -- see module docstring.)"""
    n = graph.nv()
    if n == 0:
        return []
    # maximum cardinality search -> reverse is a perfect elimination ordering
    weight = [0] * (n + 1)
    numbered = [False] * (n + 1)
    order = []
    for _ in range(n):
        v = max((u for u in range(1, n + 1) if not numbered[u]), key=lambda u: (weight[u], -u))
        numbered[v] = True
        order.append(v)
        for u in graph.adj[v - 1]:
            if not numbered[u]:
                weight[u] += 1
    pos = {v: i for i, v in enumerate(order)}
    cands = []
    for v in order:  # clique = v + neighbours numbered before v
        cands.append(frozenset([v] + [u for u in graph.adj[v - 1] if pos[u] < pos[v]]))
    cands = sorted(set(cands), key=len, reverse=True)
    maximal = []
    for c in cands:
        if not any(c < m for m in maximal):
            maximal.append(c)
    pre = lambda code: graph.vdata[graph.label_for(code)]
    maximal.sort(key=lambda c: sorted((pre(u) for u in c), reverse=True), reverse=True)
    return [sorted(c) for c in maximal]


def init_clustergraph(tag):
    return MetaGraph(tag)


def kruskal_mst(g: MetaGraph, weight, minimize=True):
    """Graphs.kruskal_mst restated: stable sort of edges (in `edges(g)` order)
    by weight, union-find."""
    el = g.edges()
    w = [weight(g.label_for(s), g.label_for(d)) for s, d in el]
    idx = sorted(range(len(el)), key=(lambda i: w[i]) if minimize else (lambda i: -w[i]))
    parent = list(range(g.nv() + 1))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    mst = []
    if g.nv() <= 1:
        return mst
    for i in idx:
        s, d = el[i]
        rs, rd = find(s), find(d)
        if rs != rd:
            parent[rs] = rd
            mst.append((s, d))
            if len(mst) >= g.nv() - 1:
                break
    return mst


def cliquetree(graph: MetaGraph, order_hint: Optional[Sequence[Sequence[int]]] = None) -> MetaGraph:
    """src/clustergraph.jl:759-820.  `order_hint`: optional list of cliques
    (each a collection of preorder indices) fixing the cluster order."""
    mc = maximal_cliques_chordal(graph)
    pre = lambda code: graph.vdata[graph.label_for(code)]
    if order_hint is not None:
        have = {frozenset(pre(u) for u in cl): cl for cl in mc}
        if order_hint and isinstance(order_hint[0], str):  # cluster labels
            bylab = {}
            for k in have:
                inds = sorted(k, reverse=True)
                names = {graph.vdata[l]: l for l in graph.labels}
                bylab["".join(names[i] for i in inds)] = k
            want = [bylab[l] for l in order_hint]
        else:
            want = [frozenset(c) for c in order_hint]
        if set(want) != set(have):
            raise ValueError("order_hint is not the set of maximal cliques")
        mc = [have[w] for w in want]
    mg = init_clustergraph("cliquetree")
    node2clique: Dict[int, List[int]] = {}
    for code, cl in enumerate(mc, start=1):
        inds = sorted((pre(u) for u in cl), reverse=True)
        by = {pre(u): graph.label_for(u) for u in cl}
        vdat = [by[i] for i in inds]
        mg.add_vertex("".join(vdat), (vdat, inds))
        for ni in inds:
            node2clique.setdefault(ni, []).append(code)
    for node in sorted(node2clique, reverse=True):
        cl = node2clique[node]
        for i1, c1 in enumerate(cl):
            l1 = mg.label_for(c1)
            for i2 in range(i1):
                l2 = mg.label_for(cl[i2])
                if mg.has_edge(l1, l2):
                    mg.edata[mg.arrange(l1, l2)].append(node)
                else:
                    mg.add_edge(l1, l2, [node])
    mst = set(kruskal_mst(mg, lambda a, b: len(mg.sepset(a, b)), minimize=False))
    for s, d in mg.edges():
        if (s, d) not in mst:
            mg.rem_edge(mg.label_for(s), mg.label_for(d))
    return mg


# --------------------------------------------------------------------------
# Bethe  (src/clustergraph.jl:473-523)
# --------------------------------------------------------------------------
def betheclustergraph(net: Network) -> MetaGraph:
    cg = init_clustergraph("Bethe")
    idx = net.preorder_index()
    prenodes = net.vec_node
    names = [n.name for n in prenodes]
    node2cluster: Dict[int, Tuple[str, List[str]]] = {}
    node2code: Dict[int, int] = {}
    code = 0
    for noi in range(len(prenodes), 0, -1):
        n = prenodes[noi - 1]
        o = sorted((idx[id(p)] for p in n.parents()), reverse=True)
        nodeind = [noi] + o
        nodesym = [names[i - 1] for i in nodeind]
        if len(nodeind) <= 1:
            continue
        sub = False
        for ch in n.children():
            ch_code = node2code[idx[id(ch)]]
            if set(nodeind) <= set(cg.vdata[cg.label_for(ch_code)][1]):
                sub = True
                node2code[noi] = ch_code
                break
        if sub:
            continue
        fname = "".join(nodesym)
        code += 1
        node2code[noi] = code
        cg.add_vertex(fname, (nodesym, nodeind))
        for nns, nni in zip(nodesym, nodeind):
            if nni in node2cluster:
                node2cluster[nni][1].append(fname)
            else:
                node2cluster[nni] = (nns, [fname])
    for ni in sorted(node2cluster, reverse=True):
        ns, cl = node2cluster[ni]
        if len(cl) <= 1:
            continue
        cg.add_vertex(ns, ([ns], [ni]))
        for lab in cl:
            cg.add_edge(ns, lab, [ni])
    return cg


# --------------------------------------------------------------------------
# LTRIP  (src/clustergraph.jl:330-344, 530-598)
# --------------------------------------------------------------------------
def ltrip_default_clusters(net: Network):
    """LTRIP(net), src/clustergraph.jl:330-333: the node families themselves, in
    preorder (so the singleton root family [1] is a cluster too)."""
    return [list(nf) for nf in nodefamilies(net)]


def ltripclustergraph(net: Network, clusters: Optional[List[List[int]]] = None) -> MetaGraph:
    if clusters is None:
        clusters = ltrip_default_clusters(net)
    fams = nodefamilies(net)
    for nf in fams:
        if not any(set(nf) <= set(c) for c in clusters):
            raise ValueError("`clusters` is not family preserving with respect to `net`")
    names = [n.name for n in net.vec_node]
    clustg = init_clustergraph("ltrip")
    aux = MetaGraph("connectionweights")
    node2cluster: Dict[int, List[int]] = {}
    for code, inds in enumerate(clusters, start=1):
        cdat = [names[i - 1] for i in inds]
        cname = "".join(cdat)
        clustg.add_vertex(cname, (cdat, list(inds)))
        aux.add_vertex(cname, (cdat, list(inds)))
        for ni in inds:
            node2cluster.setdefault(ni, []).append(code)
        for code2 in range(1, code):
            w = len(set(inds) & set(clusters[code2 - 1]))
            if w > 0:
                aux.add_edge(cname, aux.label_for(code2), w)
    for ni in sorted(node2cluster, reverse=True):
        cl = node2cluster[ni]
        sg, vmap = induced_subgraph_vertices(aux, cl)
        if sg.ne() == 0:
            continue
        maxw = max(sg.edata.values())
        cw = {lab: 0 for lab in sg.labels}
        for (a, b), w in sg.edata.items():
            if w == maxw:
                cw[a] += 1
                cw[b] += 1
        for k in list(sg.edata):
            sg.edata[k] += cw[k[0]] + cw[k[1]]
        mst = kruskal_mst(sg, lambda a, b: sg.sepset(a, b), minimize=False)
        for s, d in mst:
            l1, l2 = sg.label_for(s), sg.label_for(d)
            if clustg.has_edge(l1, l2):
                clustg.edata[clustg.arrange(l1, l2)].append(ni)
            else:
                clustg.add_edge(l1, l2, [ni])
    return clustg


# --------------------------------------------------------------------------
# Join-graph structuring  (src/clustergraph.jl:402-410, 605-736)
# --------------------------------------------------------------------------
def _assign(bucket: Dict[int, List[List[int]]], new: List[int], maxsize: int):
    """src/clustergraph.jl:705-736."""
    for sz in sorted(bucket, reverse=True):
        mbs = bucket[sz]
        for i, mb in enumerate(mbs):
            merged = sorted(set(new) | set(mb))
            if len(merged) <= maxsize:
                mbs.pop(i)
                if not mbs:
                    del bucket[sz]
                bucket.setdefault(len(merged), []).append(merged)
                return merged, mb
    bucket.setdefault(len(new), []).append(new)
    return new, []


def joingraph(net: Network, maxclustersize: int) -> MetaGraph:
    """src/clustergraph.jl:605-688.  Minibuckets of one bucket are visited by
    increasing size (the reference iterates a Dict: hash order, see module
    docstring)."""
    fams = nodefamilies(net)
    if maxclustersize < max(len(f) for f in fams):
        raise ValueError("maxclustersize is smaller than the size of largest node family")
    g = moralize(net)
    ordering = triangulate_minfill(g)
    elim2pre = [g.vdata[ns] for ns in ordering]  # elimination position -> preorder idx
    pre2elim = {p: i + 1 for i, p in enumerate(elim2pre)}
    buckets: Dict[int, Dict[int, List[List[int]]]] = {i: {} for i in range(1, len(ordering) + 1)}
    cg = init_clustergraph("auxiliary")
    for nf in fams:
        mb = sorted(pre2elim[p] for p in nf)
        _assign(buckets[mb[0]], mb, maxclustersize)

    def cluster_of(mb):
        inds = sorted((elim2pre[k - 1] for k in mb), reverse=True)
        vdat = [ordering[pre2elim[p] - 1] for p in inds]
        return "".join(vdat), vdat, inds

    for i in range(1, len(ordering) + 1):
        bd = buckets[i]
        bi = elim2pre[i - 1]
        prev = None
        for sz in list(bd.keys()):
            for mb in list(bd[sz]):
                lab, vdat, inds = cluster_of(mb)
                cg.add_vertex(lab, (vdat, inds))
                if prev is not None:
                    cg.add_edge(prev, lab, [bi])
                prev = lab
                mb_new = mb[1:]
                if not mb_new:
                    continue
                mb1, mb2 = _assign(buckets[mb_new[0]], list(mb_new), maxclustersize)
                lab1, vdat1, inds1 = cluster_of(mb1)
                cg.add_vertex(lab1, (vdat1, inds1))
                cg.add_edge(lab, lab1, [p for p in inds if p != bi])
                if len(mb1) != len(mb2) and mb2:
                    lab2, _, _ = cluster_of(mb2)
                    if cg.has_vertex(lab2):
                        for labn in cg.neighbor_labels(lab2):
                            cg.add_edge(lab1, labn, cg.sepset(lab2, labn))
                        cg.delete_vertex(lab2)
    return cg


# --------------------------------------------------------------------------
# dispatch  (src/clustergraph.jl:452-466)
# --------------------------------------------------------------------------
def clustergraph(net: Network, method: str, *, maxclustersize: int = 0, clusters=None,
                 order_hint=None, preprocess=True) -> MetaGraph:
    """clustergraph!(net, method).  method in {"cliquetree","bethe","ltrip","jgs"}."""
    if preprocess:
        preprocessnet(net)
    m = method.lower()
    if m == "cliquetree":
        g = moralize(net)
        triangulate_minfill(g)
        return cliquetree(g, order_hint)
    if m == "bethe":
        return betheclustergraph(net)
    if m == "ltrip":
        return ltripclustergraph(net, clusters)
    if m in ("jgs", "joingraph", "joingraphstructuring"):
        return joingraph(net, maxclustersize)
    raise ValueError(method)


# --------------------------------------------------------------------------
# sub-graphs
# --------------------------------------------------------------------------
def induced_subgraph_vertices(g: MetaGraph, codes: Sequence[int]):
    """Graphs.induced_subgraph(g, vlist): new code k <-> codes[k-1]."""
    sg = MetaGraph(g.tag)
    for c in codes:
        lab = g.label_for(c)
        sg.add_vertex(lab, g.vdata[lab])
    cs = set(codes)
    for c in codes:
        for nb in g.adj[c - 1]:
            if nb in cs and nb > c:
                l1, l2 = g.label_for(c), g.label_for(nb)
                d = g.edata[g.arrange(l1, l2)]
                sg.add_edge(l1, l2, list(d) if isinstance(d, list) else d)
    return sg, list(codes)


def induced_subgraph_edges(g: MetaGraph, elist: Sequence[Tuple[int, int]]):
    """Graphs.induced_subgraph(g, edgelist): vertices renumbered in order of
    first appearance in the edge list."""
    sg = MetaGraph(g.tag)
    newvid: Dict[int, int] = {}
    vmap: List[int] = []
    for u, v in elist:
        for i in (u, v):
            if i not in newvid:
                lab = g.label_for(i)
                sg.add_vertex(lab, g.vdata[lab])
                newvid[i] = sg.nv()
                vmap.append(i)
    for u, v in elist:
        l1, l2 = g.label_for(u), g.label_for(v)
        d = g.edata[g.arrange(l1, l2)]
        sg.add_edge(l1, l2, list(d) if isinstance(d, list) else d)
    return sg, vmap


def nodesubtree(cg: MetaGraph, ns: str, node_ind: Optional[int] = None):
    """src/clustergraph.jl:219-240."""
    codes = [c for c in range(1, cg.nv() + 1) if ns in cg.vdata[cg.label_for(c)][0]]
    if not codes:
        raise ValueError(f"no cluster with node labelled {ns}")
    if node_ind is None:
        d = cg.vdata[cg.label_for(codes[0])]
        node_ind = d[1][d[0].index(ns)]
    sg, vmap = induced_subgraph_vertices(cg, codes)
    for (l1, l2) in sg.edge_labels():
        if node_ind not in sg.sepset(l1, l2):
            sg.rem_edge(l1, l2)
    return sg, vmap


def check_runningintersection(cg: MetaGraph, net: Network):
    """src/clustergraph.jl:200-208."""
    res = []
    for i, n in enumerate(net.vec_node, start=1):
        sg, _ = nodesubtree(cg, n.name, i)
        res.append((n.name, is_tree(sg)))
    return res


def isfamilypreserving(clusters, net: Network):
    """src/clustergraph.jl:169-181."""
    fams = nodefamilies(net)
    inc = [[set(nf) <= set(cl) for cl in clusters] for nf in fams]
    return all(any(r) for r in inc), inc


def get_nodesymbols2index(cg: MetaGraph):
    """src/clustergraph.jl:856-860 (a Dict in the reference: iteration order
    there is hash order; here insertion order)."""
    d = {}
    for l in cg.labels:
        for ns, ni in zip(*cg.vdata[l]):
            d[ns] = ni
    return d


# --------------------------------------------------------------------------
# schedules  (src/clustergraph.jl:881-962, 1022-1053)
# --------------------------------------------------------------------------
def default_rootcluster(cg: MetaGraph, prenodes=None) -> int:
    """src/clustergraph.jl:1022-1029 (with prenodes) and :1043-1053 (without).
    Returns a 1-based cluster code."""
    if prenodes is not None:
        best, bs = None, None
        for c, lab in enumerate(cg.labels, start=1):
            nl = cg.vdata[lab][1]
            sc = sum(1 for i in nl if prenodes[i - 1].leaf) if 1 in nl else math.inf
            if bs is None or sc < bs:
                best, bs = c, sc
        return best
    i0 = min(cg.vdata[lab][1][-1] for lab in cg.labels)
    best, bs = None, None
    for c, lab in enumerate(cg.labels, start=1):
        nl = cg.vdata[lab][1]
        sc = (0 if len(nl) == 1 else nl[-2]) if i0 in nl else math.inf
        if bs is None or sc < bs:
            best, bs = c, sc
    return best


def _dfs_parents(g: MetaGraph, s: int):
    n = g.nv()
    parents = [0] * (n + 1)
    seen = [False] * (n + 1)
    S = [s]
    seen[s] = True
    parents[s] = s
    nbrs = [None] + [sorted(a) for a in g.adj]
    while S:
        v = S[-1]
        u = 0
        for w in nbrs[v]:
            if not seen[w]:
                u = w
                break
        if u == 0:
            S.pop()
        else:
            seen[u] = True
            S.append(u)
            parents[u] = v
    return parents


def _topological_sort_by_dfs(n: int, out: List[List[int]]):
    color = [0] * (n + 1)
    verts = []
    for v in range(1, n + 1):
        if color[v] != 0:
            continue
        S = [v]
        color[v] = 1
        while S:
            u = S[-1]
            w = 0
            for x in out[u]:
                if color[x] == 1:
                    raise ValueError("The input graph contains at least one loop.")
                if color[x] == 0:
                    w = x  # no break: the LAST unvisited out-neighbour is taken
            if w != 0:
                color[w] = 1
                S.append(w)
            else:
                color[u] = 2
                verts.append(u)
                S.pop()
    return verts[::-1]


def spanningtree_clusterlist(cg: MetaGraph, root=None, prenodes=None):
    """src/clustergraph.jl:881-894.  Returns (parent_labels, child_labels,
    parent_indices, child_indices), 1-based cluster codes, edges in preorder."""
    rootj = default_rootcluster(cg, prenodes) if root is None else root
    par = _dfs_parents(cg, rootj)
    n = cg.nv()
    out = [[] for _ in range(n + 1)]
    for v in range(1, n + 1):
        u = par[v]
        if u > 0 and u != v:
            out[u].append(v)
    for l in out:
        l.sort()
    topo = _topological_sort_by_dfs(n, out)
    child = topo[1:]
    parent = [par[c] for c in child]
    return ([cg.label_for(j) for j in parent], [cg.label_for(j) for j in child], parent, child)


def spanningtrees_clusterlist(cg: MetaGraph, prenodes):
    """src/clustergraph.jl:908-937."""
    w = {k: 0 for k in cg.edata}
    aux = MetaGraph("edgeweights")
    for l in cg.labels:
        aux.add_vertex(l, cg.vdata[l])
    for (l1, l2) in cg.edge_labels():
        aux.add_edge(l1, l2, 0)
    sched = []
    while any(v == 0 for v in aux.edata.values()):
        mst = kruskal_mst(aux, lambda a, b: aux.sepset(a, b), minimize=True)
        sg, vmap = induced_subgraph_edges(aux, mst)
        pl, cl, pj, cj = spanningtree_clusterlist(sg, prenodes=prenodes)
        pj = [vmap[j - 1] for j in pj]
        cj = [vmap[j - 1] for j in cj]
        sched.append((pl, cl, pj, cj))
        for s, d in mst:
            k = aux.arrange(aux.label_for(s), aux.label_for(d))
            aux.edata[k] += 1
    return sched


def nodesubtree_clusterlist(cg: MetaGraph, ns: str):
    """src/clustergraph.jl:953-962."""
    sg, vmap = nodesubtree(cg, ns)
    rootj = default_rootcluster(sg)
    pl, cl, pj, cj = spanningtree_clusterlist(sg, root=rootj)
    return (pl, cl, [vmap[j - 1] for j in pj], [vmap[j - 1] for j in cj])
