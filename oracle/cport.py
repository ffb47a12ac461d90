"""ctypes front-end of the C/OpenMP oracle twin (oracle/c/pgbp_oracle.c).

TEST INFRASTRUCTURE (see oracle/__init__.py): the fast checker at full batch
sizes and the timed CPU baseline of bench.py.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os

import numpy as np

from .beliefs import scopeindex

HERE = os.path.dirname(os.path.abspath(__file__))
i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
P = C.POINTER


class _Graph(C.Structure):
    _fields_ = [("nclusters", i32), ("nsepsets", i32), ("ntraits", i32), ("dim", P(i32)), ("off", P(i64)),
                ("sep_a", P(i32)), ("sep_b", P(i32)), ("up_off", P(i32)), ("up", P(i32)), ("roff", P(i64))]


class _Fam(C.Structure):
    _fields_ = [("nnodes", i32), ("ntips", i32), ("root_fixed", i32), ("node_cluster", P(i32)), ("mem_off", P(i32)),
                ("mem_pos", P(i32)), ("mem_length", P(f64)), ("mem_gamma", P(f64)), ("mem_color", P(i32)),
                ("node_datarow", P(i32))]


_dll = {}


def dll(quad=False):
    """quad=True: the same C source compiled in IEEE binary128 (oracle/c/build.py) -- the >= 100-bit adjudicator."""
    if quad not in _dll:
        spec = importlib.util.spec_from_file_location("pgbp_oracle_build", os.path.join(HERE, "c", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _dll[quad] = C.CDLL(mod.build(quad=quad))
        _dll[quad].pgbpo_num_threads.restype = i32
    return _dll[quad]


def _i(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).ravel())


def _p(a, t):
    return a.ctypes.data_as(P(t))


class COracle:
    """dims/sepsets/upind/trees/families in the same (0-based) form the product's
    ClusterGraphPlan takes; built here from oracle-side beliefs."""

    def __init__(self, beliefs, nclusters, cluster_labels, schedule, families, ntraits):
        lab2idx = {l: k for k, l in enumerate(cluster_labels)}
        sc, up = [], []
        for s in beliefs[nclusters:]:
            a, b = lab2idx[s.metadata[0]], lab2idx[s.metadata[1]]
            sc.append((a, b))
            up.append((scopeindex(s, beliefs[a]), scopeindex(s, beliefs[b])))
        trees = [([j - 1 for j in spt[2]], [j - 1 for j in spt[3]]) for spt in schedule]
        self._init([b.dimension() for b in beliefs], nclusters, sc, up, trees, families, ntraits)

    @classmethod
    def from_plan_dict(cls, d):
        """From a workloads/*.json plan dump (same arrays the product's plan takes)."""
        self = cls.__new__(cls)
        self._init(d["belief_dim"], d["nclusters"], d["sepset_clusters"], d["upind"], d["trees"], d["families"],
                   d["ntraits"])
        return self

    def _init(self, dims, nclusters, sepset_clusters, upind, trees, families, ntraits):
        self.nclusters, self.nsepsets = nclusters, len(dims) - nclusters
        self.dim = _i(dims)
        off = [0]
        for m in self.dim:
            off.append(off[-1] + m * m + m + 1)
        self.off = np.array(off, dtype=np.int64)
        sa, sb, upo, up, roff = [], [], [0], [], [0]
        for j, ((a, b), (ua, ub)) in enumerate(zip(sepset_clusters, upind)):
            sa.append(a); sb.append(b)
            up.extend(int(x) for x in ua); upo.append(len(up))
            up.extend(int(x) for x in ub); upo.append(len(up))
            d = int(self.dim[nclusters + j])
            roff.append(roff[-1] + d * d + d); roff.append(roff[-1] + d * d + d)
        self.sep_a, self.sep_b, self.up_off, self.up = _i(sa or [0]), _i(sb or [0]), _i(upo), _i(up or [0])
        self.roff = np.array(roff, dtype=np.int64)
        self.G = _Graph(nclusters, self.nsepsets, ntraits, _p(self.dim, i32), _p(self.off, i64), _p(self.sep_a, i32),
                        _p(self.sep_b, i32), _p(self.up_off, i32), _p(self.up, i32), _p(self.roff, i64))
        sepof = {frozenset((a, b)): j for j, (a, b) in enumerate(zip(sa, sb))}
        toff, tsep, tpar, tchi = [0], [], [], []
        for par, chi in trees:
            for pj, cj in zip(par, chi):
                tpar.append(pj); tchi.append(cj); tsep.append(sepof[frozenset((pj, cj))])
            toff.append(len(tpar))
        self.tree_off, self.tsep, self.tpar, self.tchi = _i(toff), _i(tsep or [0]), _i(tpar or [0]), _i(tchi or [0])
        schedule = trees
        self.ntrees = len(schedule)
        f = families
        self._fam = [_i(f["node_cluster"]), _i(f["mem_off"]), _i(f["mem_pos"]),
                     np.ascontiguousarray(f["mem_length"], dtype=float), np.ascontiguousarray(f["mem_gamma"], dtype=float),
                     _i(f["mem_color"]), _i(f["node_datarow"])]
        a = self._fam
        self.F = _Fam(f["nnodes"], f["ntips"], f["root_fixed"], _p(a[0], i32), _p(a[1], i32), _p(a[2], i32),
                      _p(a[3], f64), _p(a[4], f64), _p(a[5], i32), _p(a[6], i32))
        self.state_size = int(self.off[-1])

    def run_batch(self, params, tipdata, ncolors=1, pairing="zip", niter=1, post=True, pre=True, residnorm=True,
                  auto=False, root_belief=0, want_fe=False, want_state=False, nthreads=0, B=None, reg_bycluster=False,
                  quad=False):
        params = np.ascontiguousarray(np.atleast_2d(np.asarray(params, dtype=float)))
        tip = np.ascontiguousarray(np.asarray(tipdata, dtype=float))
        if tip.ndim == 2:
            tip = tip[None]
        npar, nd = params.shape[0], tip.shape[0]
        if B is None:
            B = npar * nd if pairing == "product" else max(npar, nd)
        ll = np.empty(B); st = np.zeros(B, dtype=np.int32); isc = np.zeros(B, dtype=np.int32)
        fe = np.empty((B, 3)) if want_fe else None
        so = np.empty((B, self.state_size)) if want_state else None
        rc = dll(quad).pgbpo_run_batch(C.byref(self.G), C.byref(self.F), i32(ncolors), _p(params, f64), i64(npar),
                                   _p(tip, f64), i64(nd), i32(1 if pairing == "product" else 0), i32(self.ntrees),
                                   _p(self.tree_off, i32), _p(self.tsep, i32), _p(self.tpar, i32), _p(self.tchi, i32),
                                   i32(niter), i32(post), i32(pre), i32(residnorm), i32(auto), i32(root_belief), i64(B),
                                   _p(ll, f64), _p(st, i32), None if fe is None else _p(fe, f64),
                                   None if so is None else _p(so, f64), _p(isc, i32), i32(nthreads), i32(int(reg_bycluster)))
        assert rc == 0
        out = dict(loglik=ll, status=st, iscal=isc.astype(bool))
        if want_fe:
            out["fe"] = fe
        if want_state:
            out["state"] = so
        return out

    def unpack(self, state_row, b):
        """(J, h, g) of belief b (0-based) from one row of `state`."""
        m = int(self.dim[b]); o = int(self.off[b])
        J = state_row[o:o + m * m].reshape(m, m).T
        return J, state_row[o + m * m:o + m * m + m], state_row[o + m * m + m]
