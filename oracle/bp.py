"""Message passing, calibration, scores, regularisation -- restates
src/beliefupdates.jl:483-488,579-665, src/calibration.jl:35-161,
src/clustergraphbeliefs.jl:26-403, src/score.jl:11-182,
src/beliefs.jl:994-1075 for canonical beliefs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  One ClusterGraphBelief = one
data set x one parameter vector, sequential, exactly like the reference.
"""
from __future__ import annotations

import math

import numpy as np

from .beliefs import (CLUSTER, SEPSET, CanonicalBelief, ClusterFactor, MessageResidual,
                      scopeindex, scopeindex_node)
from .canonical import (EPS, LOG2PI, BPPosDefException, _solve_u, _solve_ut, chol_upper,
                        integratebelief, marginalize)
from .clustergraph import (get_nodesymbols2index, nodesubtree, spanningtree_clusterlist)


class ClusterGraphBelief:
    """src/clustergraphbeliefs.jl:26-53, 89-116.  Indices are 1-based like the
    reference; `belief` is a Python list (0-based storage)."""

    def __init__(self, beliefs, node2cluster, node2family, node2fixed, cluster2nodes):
        i = next((k for k, b in enumerate(beliefs) if b.type == SEPSET), None)
        nc = len(beliefs) if i is None else i
        if not all(b.type == CLUSTER for b in beliefs[:nc]):
            raise ValueError("clusters are not consecutive")
        if not all(b.type == SEPSET for b in beliefs[nc:]):
            raise ValueError("sepsets are not consecutive")
        self.belief = beliefs
        self.nclusters = nc
        self.cdict = {beliefs[j].metadata: j + 1 for j in range(nc)}
        self.sdict = {frozenset(beliefs[j].metadata): j + 1 for j in range(nc, len(beliefs))}
        self.messageresidual = {}
        for j in range(nc, len(beliefs)):
            s = beliefs[j]
            l1, l2 = s.metadata
            self.messageresidual[(l1, l2)] = MessageResidual(s.J, s.h)
            self.messageresidual[(l2, l1)] = MessageResidual(s.J, s.h)
        self.factor = [ClusterFactor(b) for b in beliefs[:nc]]
        self.node2cluster, self.node2family = node2cluster, node2family
        self.node2fixed, self.cluster2nodes = node2fixed, cluster2nodes

    def nsepsets(self):
        return len(self.belief) - self.nclusters

    def clusterindex(self, lab):
        return self.cdict[lab]

    def sepsetindex(self, l1, l2):
        return self.sdict[frozenset((l1, l2))]


def init_factors_frombeliefs(cgb: ClusterGraphBelief):
    """src/beliefs.jl:747-761."""
    for fa, be in zip(cgb.factor, cgb.belief):
        fa.h[:] = be.h
        fa.J[:] = be.J
        fa.g = be.g


def init_beliefs_reset_fromfactors(cgb: ClusterGraphBelief):
    """src/clustergraphbeliefs.jl:126-139."""
    for i in range(cgb.nclusters):
        b, f = cgb.belief[i], cgb.factor[i]
        b.h[:] = f.h
        b.J[:] = f.J
        b.g = f.g
    for b in cgb.belief[cgb.nclusters:]:
        b.h[:] = 0
        b.J[:] = 0
        b.g = 0.0


def init_messagecalibrationflags_reset(cgb: ClusterGraphBelief, reset_kl=True):
    """src/clustergraphbeliefs.jl:146-150."""
    for m in cgb.messageresidual.values():
        m.reset_flags(reset_kl)


# --------------------------------------------------------------------------
# one message (src/beliefupdates.jl:634-665)
# --------------------------------------------------------------------------
def divide(sepset: CanonicalBelief, h, J, g):
    """src/beliefupdates.jl:579-587."""
    dh, dJ, dg = h - sepset.h, J - sepset.J, g - sepset.g
    sepset.h[:] = h
    sepset.J[:] = J
    sepset.g = g
    return dh, dJ, dg


def mult(cluster_to: CanonicalBelief, upind, dh, dJ, dg):
    """src/beliefupdates.jl:483-488."""
    cluster_to.h[upind] += dh
    cluster_to.J[np.ix_(upind, upind)] += dJ
    cluster_to.g += dg


def propagate_belief(cluster_to, sepset, cluster_from, residual=None):
    """Returns None, or the BPPosDefException (returned, not raised:
    src/beliefupdates.jl:640-644)."""
    try:
        keep = scopeindex(sepset, cluster_from)
        h, J, g = marginalize(cluster_from.h, cluster_from.J, cluster_from.g, keep,
                              metadata=cluster_from.metadata)
    except BPPosDefException as ex:
        return ex
    dh, dJ, dg = divide(sepset, np.array(h, dtype=float), np.array(J, dtype=float), g)
    mult(cluster_to, scopeindex(sepset, cluster_to), dh, dJ, dg)
    if residual is not None:
        residual.dh[:] = dh
        residual.dJ[:] = dJ
    return None


def iscalibrated_residnorm_update(res: MessageResidual, atol=1e-5):
    """src/beliefs.jl:994-1003 (p = Inf)."""
    def nrm(x):
        return 0.0 if x.size == 0 else float(np.max(np.abs(x / math.sqrt(x.size))))
    res.iscalibrated_resid = (nrm(res.dh) <= atol) and (nrm(res.dJ) <= atol)
    return res.iscalibrated_resid


def residual_kldiv_update(res: MessageResidual, sepset: CanonicalBelief, atol=1e-5):
    """src/beliefs.jl:1060-1075."""
    if sepset.J.size == 0:
        return True
    try:
        U0 = chol_upper(sepset.J)
        mu0 = _solve_u(U0, _solve_ut(U0, sepset.h))
        sepset.mu[:] = mu0
        J1 = sepset.J - res.dJ
        U1 = chol_upper(J1)
        mu1 = _solve_u(U1, _solve_ut(U1, sepset.h - res.dh))
    except BPPosDefException:
        return False
    J0inv_dJ = _solve_u(U0, _solve_ut(U0, res.dJ))
    d = mu1 - mu0
    ld0 = 2 * float(np.sum(np.log(np.diag(U0))))
    ld1 = 2 * float(np.sum(np.log(np.diag(U1))))
    res.kldiv = (-float(np.trace(J0inv_dJ)) + float(d @ J1 @ d) + ld0 - ld1) / 2
    res.iscalibrated_kl = abs(res.kldiv) <= atol
    return res.iscalibrated_kl


# --------------------------------------------------------------------------
# traversals and calibrate! (src/calibration.jl:35-161)
# --------------------------------------------------------------------------
def propagate_1traversal_postorder(cgb, pa_lab, ch_lab, pa_j, ch_j, verbose=True,
                                   update_residualnorm=True, update_residualkldiv=False):
    b, mr = cgb.belief, cgb.messageresidual
    for i in range(len(pa_lab) - 1, -1, -1):
        sepset = b[cgb.sepsetindex(pa_lab[i], ch_lab[i]) - 1]
        mrss = mr[(pa_lab[i], ch_lab[i])]
        flag = propagate_belief(b[pa_j[i] - 1], sepset, b[ch_j[i] - 1], mrss)
        if flag is None:
            if update_residualnorm:
                iscalibrated_residnorm_update(mrss)
            if update_residualkldiv:
                residual_kldiv_update(mrss, sepset)
        else:
            return False
    return True


def propagate_1traversal_preorder(cgb, pa_lab, ch_lab, pa_j, ch_j, verbose=True,
                                  update_residualnorm=True, update_residualkldiv=False):
    b, mr = cgb.belief, cgb.messageresidual
    for i in range(len(pa_lab)):
        sepset = b[cgb.sepsetindex(pa_lab[i], ch_lab[i]) - 1]
        mrss = mr[(ch_lab[i], pa_lab[i])]
        flag = propagate_belief(b[ch_j[i] - 1], sepset, b[pa_j[i] - 1], mrss)
        if flag is None:
            if update_residualnorm:
                iscalibrated_residnorm_update(mrss)
            if update_residualkldiv:
                residual_kldiv_update(mrss, sepset)
        else:
            return False
    return True


def iscalibrated_residnorm(cgb):
    """src/clustergraphbeliefs.jl:168-169."""
    return all(m.iscalibrated_resid for m in cgb.messageresidual.values())


def iscalibrated_kl(cgb):
    return all(m.iscalibrated_kl for m in cgb.messageresidual.values())


def calibrate_tree(cgb, spt, verbose=True, up_resnorm=True, up_reskldiv=False):
    """src/calibration.jl:72-84: both traversals always run."""
    pos = propagate_1traversal_postorder(cgb, *spt, verbose, up_resnorm, up_reskldiv)
    pre = propagate_1traversal_preorder(cgb, *spt, verbose, up_resnorm, up_reskldiv)
    if not (pos and pre):
        return (False, False)
    return (True, iscalibrated_residnorm(cgb))


def calibrate(cgb, schedule, niter=1, auto=False, info=False, verbose=True,
              update_residualnorm=True, update_residualkldiv=False):
    """src/calibration.jl:35-60 -> (succ, iscal) [+ where calibration was first
    reached, as `calibrate.last_info`]."""
    succ, iscal = False, False
    calibrate.last_info = None
    for i in range(1, niter + 1):
        for j, spt in enumerate(schedule, start=1):
            succ, iscal = calibrate_tree(cgb, spt, verbose, update_residualnorm, update_residualkldiv)
            if not succ:
                return succ, iscal
            if iscal:
                if calibrate.last_info is None:
                    calibrate.last_info = (i, j)
                if auto:
                    return succ, iscal
    return succ, iscal


# --------------------------------------------------------------------------
# integrate (src/clustergraphbeliefs.jl:190-202, src/beliefupdates.jl:168-172)
# --------------------------------------------------------------------------
def integratebelief_inplace(b: CanonicalBelief):
    mu, norm = integratebelief(b.h, b.J, b.g)
    b.mu[:] = mu
    return mu, norm


def default_sepset1(cgb):
    for j in range(cgb.nclusters, len(cgb.belief)):
        if len(cgb.belief[j].nodelabel) == 1:
            return j + 1
    raise ValueError("no sepset with a single node")


def integratebelief_cgb(cgb, j=None):
    if j is None:
        j = default_sepset1(cgb)
    return integratebelief_inplace(cgb.belief[j - 1])


# --------------------------------------------------------------------------
# scores (src/score.jl)
# --------------------------------------------------------------------------
def entropy_chol(U):
    """src/score.jl:58-62."""
    n = U.shape[0]
    if n == 0:
        return 0.0
    return (n * (LOG2PI + 1) - 2 * float(np.sum(np.log(np.diag(U))))) / 2


def entropy_matrix(J):
    """src/score.jl:63-67: logdet(Symmetric(J)) (LU in the reference)."""
    n = J.shape[0]
    if n == 0:
        return 0.0
    Js = np.triu(J) + np.triu(J, 1).T
    sign, ld = np.linalg.slogdet(Js)
    if sign == 0:
        ld = -math.inf
    elif sign < 0:
        ld = math.nan  # Julia: DomainError
    return (n * (LOG2PI + 1) - ld) / 2


def average_energy(U, mu, Jt, ht, gt):
    """src/score.jl:114-117."""
    if Jt.size == 0:
        return -gt
    JinvJt = _solve_u(U, _solve_ut(U, Jt))
    return (float(np.trace(JinvJt)) + float(mu @ Jt @ mu)) / 2 - float(ht @ mu) - gt


def free_energy(cgb):
    """src/score.jl:162-182 -> (average energy, approx entropy, free energy)."""
    ave, ent = 0.0, 0.0
    for i in range(cgb.nclusters):
        fac = cgb.factor[i]
        if fac.J.size == 0:
            ave -= fac.g
        else:
            b = cgb.belief[i]
            U = chol_upper(b.J)
            mu = _solve_u(U, _solve_ut(U, b.h))
            b.mu[:] = mu
            ave += average_energy(U, mu, fac.J, fac.h, fac.g)
            ent += entropy_chol(U)
    for b in cgb.belief[cgb.nclusters:]:
        ent -= entropy_matrix(b.J)
    return ave, ent, ave - ent


def factored_energy(cgb):
    """src/score.jl:151-154."""
    r = free_energy(cgb)
    return r[0], r[1], -r[2]


# --------------------------------------------------------------------------
# regularisation (src/clustergraphbeliefs.jl:235-403)
# --------------------------------------------------------------------------
def regularizebeliefs_1clustersepset(cluster, sepset, eps):
    """:264-275."""
    upind = scopeindex(sepset, cluster)
    if upind.size == 0:
        return
    cluster.J[upind, upind] += eps
    d = np.arange(sepset.J.shape[0])
    sepset.J[d, d] += eps


def regularizebeliefs_bycluster(cgb, cgraph, clusterlab=None):
    """:235-249."""
    labs = cgraph.labels if clusterlab is None else [clusterlab]
    for lab in labs:
        cl = cgb.belief[cgb.clusterindex(lab) - 1]
        eps = max(EPS, float(np.max(np.abs(cl.J))) if cl.J.size else 0.0)
        for nb in cgraph.neighbor_labels(lab):
            ss = cgb.belief[cgb.sepsetindex(lab, nb) - 1]
            regularizebeliefs_1clustersepset(cl, ss, eps)


def bynodesubtree_program(cgb, cgraph, node_order=None):
    """The index program regularizebeliefs_bynodesubtree! executes
    (:306-340), one entry per network node with a non-trivial cluster subtree:
    (cluster indices whose |J| max defines eps,
     [(child cluster idx, sepset idx, c_ind, s_ind), ...]) -- 1-based belief
    indices, 0-based variable positions.  The reference walks a Dict (hash
    order); here nodes are taken in `node_order` (default: first appearance in
    the cluster list).  Only the speed of convergence depends on it."""
    n2i = get_nodesymbols2index(cgraph)
    order = list(n2i.items()) if node_order is None else [(ns, n2i[ns]) for ns in node_order]
    prog = []
    for ns, ni in order:
        sg, _ = nodesubtree(cgraph, ns, ni)
        if sg.nv() <= 1:
            continue
        # cluster with the largest first (= largest preorder) node index
        rootj = max(range(1, sg.nv() + 1), key=lambda c: (sg.vdata[sg.label_for(c)][1][0], -c))
        pl, cl, _, _ = spanningtree_clusterlist(sg, root=rootj)
        epscl = [cgb.clusterindex(l) for l in sg.labels]
        steps = []
        for par_l, chi_l in zip(pl, cl):
            ci = cgb.clusterindex(chi_l)
            si = cgb.sepsetindex(par_l, chi_l)
            s_ind, c_ind = scopeindex_node(ni, cgb.belief[si - 1], cgb.belief[ci - 1])
            steps.append((ci, si, c_ind, s_ind))
        prog.append((epscl, steps))
    return prog


def regularizebeliefs_bynodesubtree(cgb, cgraph, node_order=None):
    """:306-340."""
    for epscl, steps in bynodesubtree_program(cgb, cgraph, node_order):
        eps = EPS
        for ci in epscl:
            J = cgb.belief[ci - 1].J
            if J.size:
                eps = max(eps, float(np.max(np.abs(J))))
        for ci, si, c_ind, s_ind in steps:
            cgb.belief[ci - 1].J[c_ind, c_ind] += eps
            cgb.belief[si - 1].J[s_ind, s_ind] += eps


def onschedule_program(cgb, cgraph):
    """The op list regularizebeliefs_onschedule! executes (:376-403):
    ("eps", cluster idx) | ("reg", cluster idx, sepset idx) |
    ("msg", to idx, sepset idx, from idx) -- 1-based belief indices."""
    sent = set()
    prog = []
    for lab in cgraph.labels:
        ci = cgb.clusterindex(lab)
        prog.append(("eps", ci))
        tosend = []
        for nb in cgraph.neighbor_labels(lab):
            nbi = cgb.clusterindex(nb)
            ssi = cgb.sepsetindex(lab, nb)
            if (nb, lab) not in sent:
                prog.append(("reg", ci, ssi))
                sent.add((nb, lab))
            if (lab, nb) not in sent:
                tosend.append((nb, nbi, ssi))
                sent.add((lab, nb))
        for nb, nbi, ssi in tosend:
            prog.append(("msg", nbi, ssi, ci))
    return prog


def regularizebeliefs_onschedule(cgb, cgraph):
    """:376-403."""
    eps0 = math.sqrt(EPS)
    b = cgb.belief
    eps = None
    for op in onschedule_program(cgb, cgraph):
        if op[0] == "eps":
            J = b[op[1] - 1].J
            eps = max(float(np.max(np.abs(J))) if J.size else 0.0, eps0)
        elif op[0] == "reg":
            regularizebeliefs_1clustersepset(b[op[1] - 1], b[op[2] - 1], eps)
        else:
            _, to, ss, fr = op
            res = cgb.messageresidual[(b[to - 1].metadata, b[fr - 1].metadata)]
            propagate_belief(b[to - 1], b[ss - 1], b[fr - 1], res)
