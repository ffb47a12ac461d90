"""Independent dense multivariate-normal likelihood on a network.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the cross-check the
reference's own tests use in their comments (`vcv(net)` + `MvNormal`,
test/test_evomodels.jl:265-316, test/test_calibration.jl:119-124,167-175): the
joint covariance of all node states is built by the network recursion

    Cov(X_i, X_j) = sum_k gamma_k Cov(X_pa_k, X_j)                (j before i)
    Var(X_i)      = sum_k gamma_k^2 t_k R_k + sum_kl gamma_k gamma_l Cov(X_pa_k, X_pa_l)

and the tip block is factorised densely.  It shares no code with the belief
propagation path.
"""
from __future__ import annotations

import math

import numpy as np

LOG2PI = math.log(2 * math.pi)


def network_covariance(net, rate_of_edge, p, rootvar=None):
    """(n*p x n*p) covariance of all node states, node-major / trait-minor, in
    preorder.  rate_of_edge(edge) -> p x p variance rate."""
    pre = net.vec_node
    n = len(pre)
    idx = {id(v): i for i, v in enumerate(pre)}
    C = np.zeros((n * p, n * p))
    if rootvar is not None:
        C[:p, :p] = rootvar
    for i in range(1, n):
        v = pre[i]
        pes = v.parent_edges()
        sl = slice(i * p, (i + 1) * p)
        for j in range(i):
            sj = slice(j * p, (j + 1) * p)
            blk = np.zeros((p, p))
            for e in pes:
                k = idx[id(e.parent)]
                blk += e.gamma * C[k * p:(k + 1) * p, sj]
            C[sl, sj] = blk
            C[sj, sl] = blk.T
        var = np.zeros((p, p))
        for e in pes:
            var += e.gamma ** 2 * e.length * rate_of_edge(e)
            k = idx[id(e.parent)]
            for e2 in pes:
                l = idx[id(e2.parent)]
                var += e.gamma * e2.gamma * C[k * p:(k + 1) * p, l * p:(l + 1) * p]
        C[sl, sl] = var
    return C


def loglik_bm(net, tbl, taxa, rate_of_edge, mu, rootvar=None, improper=False):
    """Log-likelihood of tip data `tbl` (ntaxa x p, NaN = missing).
    rootvar=None/0 => fixed root at mu; improper=True => flat prior on the root
    mean (restricted likelihood, what an infinite root variance gives)."""
    tbl = np.asarray(tbl, dtype=float)
    p = tbl.shape[1]
    mu = np.atleast_1d(np.asarray(mu, dtype=float))
    C = network_covariance(net, rate_of_edge, p, None if improper else rootvar)
    idx = {v.name: i for i, v in enumerate(net.vec_node)}
    sel, y, tr = [], [], []
    for r, name in enumerate(taxa):
        for t in range(p):
            if not math.isnan(tbl[r, t]):
                sel.append(idx[name] * p + t)
                y.append(tbl[r, t])
                tr.append(t)
    sel = np.array(sel)
    y = np.array(y)
    S = C[np.ix_(sel, sel)]
    L = np.linalg.cholesky(S)
    ld = 2 * float(np.sum(np.log(np.diag(L))))
    n = y.size
    if not improper:
        r = np.linalg.solve(L, y - mu[tr])
        return -(n * LOG2PI + ld + float(r @ r)) / 2
    X = np.zeros((n, p))
    X[np.arange(n), tr] = 1.0
    LX = np.linalg.solve(L, X)
    Ly = np.linalg.solve(L, y)
    A = LX.T @ LX
    beta = np.linalg.solve(A, LX.T @ Ly)
    r = Ly - LX @ beta
    _, ldA = np.linalg.slogdet(A)
    return -((n - p) * LOG2PI + ld + ldA + float(r @ r)) / 2
