"""Extended-Newick reader and PhyloNetworks-style node pre-ordering.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference delegates this to
PhyloNetworks.jl (`readnewick`, `preorder!`, `nameinternalnodes!`, called from
src/clustergraph.jl:18-21), which is not vendored under /root/reference.  The
behaviour restated here is pinned by the reference's tests:
  * node order  i1,i2,C,i4,H5,i6,B2,B1,A   (test/test_evomodels.jl:156)
  * belief node labels                    (test/test_canonicalform.jl:54-56)
  * elimination order with I1..I5 names   (test/test_clustergraph.jl:11-12)
  * lazaridis cluster labels / Int8[17,16,10] (docs/src/man/getting_started.md:107-163)

Conventions restated:
  * nodes and edges are created in the order their Newick sub-string is
    *closed* (post-order of the string); a node's edge list holds the edges to
    its children in Newick order, then its parent edge(s);
  * a root of degree 1 is removed (its only child becomes the root);
  * a hybrid node `#Hx` occurs twice; one gamma missing => 1 - the other;
  * preorder: LIFO stack, children pushed in edge order (so the last child is
    visited first); a hybrid child is pushed once all its parents are visited;
  * unnamed internal nodes get names prefix+counter in net.node order.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass(eq=False)
class Edge:
    number: int
    length: float = -1.0
    gamma: float = 1.0
    hybrid: bool = False
    child: Optional["Node"] = None
    parent: Optional["Node"] = None
    _gamma_given: bool = False


@dataclass(eq=False)
class Node:
    name: str = ""
    leaf: bool = False
    hybrid: bool = False
    edges: List[Edge] = field(default_factory=list)

    def parent_edges(self):
        return [e for e in self.edges if e.child is self]

    def child_edges(self):
        return [e for e in self.edges if e.parent is self]

    def parents(self):
        return [e.parent for e in self.parent_edges()]

    def children(self):
        return [e.child for e in self.child_edges()]


class Network:
    def __init__(self):
        self.nodes: List[Node] = []
        self.edges: List[Edge] = []
        self.root: Optional[Node] = None
        self.vec_node: List[Node] = []  # preorder

    # convenience -----------------------------------------------------------
    @property
    def numnodes(self):
        return len(self.nodes)

    def tiplabels(self):
        """Leaf names in net.node order (PhyloNetworks `tiplabels`)."""
        return [n.name for n in self.nodes if n.leaf]

    def preorder_index(self):
        """dict node -> 1-based preorder index."""
        return {id(n): i + 1 for i, n in enumerate(self.vec_node)}


class _Parser:
    def __init__(self, s: str):
        self.s = s.strip()
        self.i = 0
        self.net = Network()
        self.hybrids = {}  # name -> Node

    def peek(self):
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1
        return self.s[self.i] if self.i < len(self.s) else ""

    def take(self):
        c = self.peek()
        self.i += 1
        return c

    _name_re = re.compile(r"[^\s():,;\[\]]+")

    def read_name(self):
        self.peek()
        m = self._name_re.match(self.s, self.i)
        if not m:
            return ""
        self.i = m.end()
        return m.group(0)

    def read_number(self):
        self.peek()
        m = re.compile(r"[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)").match(self.s, self.i)
        if not m:
            return None
        self.i = m.end()
        return float(m.group(0))

    def new_edge(self, child, parent, hybrid=False):
        e = Edge(number=len(self.net.edges) + 1, hybrid=hybrid, child=child, parent=parent)
        self.net.edges.append(e)
        # order mirrors setEdge!(n,e); setEdge!(parent,e)
        child.edges.append(e)
        if parent is not None:
            parent.edges.append(e)
        return e

    def subtree(self, parent: Optional[Node]):
        c = self.peek()
        if c == "(":
            n = Node()
            while True:
                self.take()  # '(' or ','
                self.subtree(n)
                c = self.peek()
                if c == ")":
                    self.take()
                    break
                if c != ",":
                    raise ValueError(f"expected , or ) at {self.i}: {self.s[self.i:self.i+20]!r}")
            name = self.read_name()
        else:
            n = Node(leaf=True)
            name = self.read_name()
            if not name:
                raise ValueError(f"leaf without a name at {self.i}")
        is_hyb = name.startswith("#")
        if parent is None:  # root
            n.name = name
            self.net.nodes.append(n)
            self.net.root = n
            return None
        if is_hyb:
            hname = name[1:]
            if hname in self.hybrids:
                other = self.hybrids[hname]
                if n.leaf:  # this occurrence is the bare reference: keep `other`
                    keep = other
                else:  # this occurrence carries the subtree: it replaces the stub
                    keep = n
                    keep.name, keep.hybrid, keep.leaf = hname, True, False
                    for e in other.edges:  # re-attach the stub's parent edge
                        e.child = keep
                    # stub's parent edges come first in time, but `keep` lists its
                    # children first (they were attached while reading its subtree)
                    keep.edges.extend(other.edges)
                    idx = self.net.nodes.index(other)
                    self.net.nodes[idx] = keep
                    self.hybrids[hname] = keep
                e = self.new_edge(keep, parent, hybrid=True)
            else:
                n.name, n.hybrid = hname, True
                # a bare `#H` reference seen first is a stub: not a leaf of the network
                n._stub = n.leaf
                n.leaf = False
                self.hybrids[hname] = n
                self.net.nodes.append(n)
                e = self.new_edge(n, parent, hybrid=True)
        else:
            n.name = name
            self.net.nodes.append(n)
            e = self.new_edge(n, parent)
        # edge data  :length:support:gamma
        if self.peek() == ":":
            self.take()
            v = self.read_number()
            if v is not None:
                e.length = v
            if self.peek() == ":":
                self.take()
                self.read_number()  # bootstrap support, ignored
                if self.peek() == ":":
                    self.take()
                    g = self.read_number()
                    if g is not None:
                        e.gamma = g
                        e._gamma_given = True
        return e

    def parse(self):
        self.subtree(None)
        # optional root edge data is ignored, then ';'
        if self.peek() == ":":
            self.take()
            self.read_number()
        net = self.net
        # hybrid gammas
        for h in self.hybrids.values():
            pe = h.parent_edges()
            given = [e for e in pe if e._gamma_given]
            if len(pe) == 2 and len(given) == 1:
                other = pe[0] if pe[1] is given[0] else pe[1]
                other.gamma = 1.0 - given[0].gamma
            elif len(given) == 0 and len(pe) >= 2:
                for e in pe:
                    e.gamma = 1.0 / len(pe)
        # remove a degree-1 root
        r = net.root
        while len(r.edges) == 1 and not r.leaf:
            e = r.edges[0]
            ch = e.child
            ch.edges.remove(e)
            net.edges.remove(e)
            net.nodes.remove(r)
            net.root = r = ch
        for k, e in enumerate(net.edges):
            e.number = k + 1
        return net


def readnewick(s: str) -> Network:
    """Parse an extended Newick string (PhyloNetworks `readnewick` restated)."""
    return _Parser(s).parse()


def preorder(net: Network) -> None:
    """PhyloNetworks `preorder!` restated (see module docstring)."""
    vec, visited = [], set()
    stack = [net.root]
    while stack:
        curr = stack.pop()
        if id(curr) in visited:
            continue
        visited.add(id(curr))
        vec.append(curr)
        for e in curr.edges:
            if e.parent is curr:
                ch = e.child
                if not e.hybrid:
                    stack.append(ch)
                elif all(id(p) in visited for p in ch.parents()):
                    stack.append(ch)
    if len(vec) != len(net.nodes):
        raise ValueError("preorder did not reach every node (is the network rooted / acyclic?)")
    net.vec_node = vec


def nameinternalnodes(net: Network, prefix: str = "I") -> None:
    """PhyloNetworks `nameinternalnodes!` restated: unnamed internal nodes get
    prefix+k, k continuing after the largest existing `prefix<int>` name."""
    rx = re.compile(r"^" + re.escape(prefix) + r"(\d+)$")
    mx = 0
    for n in net.nodes:
        m = rx.match(n.name)
        if m:
            mx = max(mx, int(m.group(1)))
    for n in net.nodes:
        if n.leaf:
            continue
        if n.name == "":
            mx += 1
            n.name = f"{prefix}{mx}"


def preprocessnet(net: Network, prefix: str = "I") -> None:
    """src/clustergraph.jl:18-21."""
    preorder(net)
    nameinternalnodes(net, prefix)


def nodefamilies(net: Network):
    """src/clustergraph.jl:136-146 -- [child, parents sorted decreasing], 1-based
    preorder indices."""
    idx = net.preorder_index()
    fam = []
    for code, n in enumerate(net.vec_node, start=1):
        o = sorted((idx[id(p)] for p in n.parents()), reverse=True)
        fam.append([code] + o)
    return fam
