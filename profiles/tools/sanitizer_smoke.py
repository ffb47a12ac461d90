"""Small C2 / C3 / C4-shape invocations of the hot path for compute-sanitizer (memcheck, racecheck):
    compute-sanitizer --tool memcheck python profiles/tools/sanitizer_smoke.py
Sizes are small (the sanitizer runs kernels 10-100x slower); every kernel family of the default dispatch is hit:
K1 (fast + scoped), register kernels, copy kernel, tile-walk, shared-memory kernels, integrate, energy, regularise,
the peer-window gather (one rank)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pgbp_b200  # noqa: E402
from pgbp_b200 import sharding  # noqa: E402

lib = pgbp_b200.default_library()


def plan_of(w):
    d = w.d
    return pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)


# C2: clique tree, residuals, lazy sepset zero, pipelined chunks off (small batch), CUDA graph on the third call
w = bench.C2()
params, tips = w.inputs(200, 0)
bt = pgbp_b200.BatchedClusterGraphBelief(plan_of(w), 200)
for _ in range(3):
    bt.assignfactors(params, tips)
    succ, iscal = bt.calibrate(None, 1)
ll = bt.integratebelief(w.d["root_cluster"] + 1)[1]
fe = bt.factored_energy()
assert succ.all() and np.allclose(ll, fe[:, 2], rtol=1e-9)
comm = sharding.PeerGather(lib, 0, 0, 1, 224)
comm.integrate_gather(bt, w.d["root_cluster"] + 1, 0)
comm.wait(bt, 0)
comm.check(bt)
assert np.array_equal(comm.read(bt, 0)[0, :200], ll)
comm.close()
print("c2 ok", ll[:2])

# C3: loopy, tile-walk kernel + regulariser + reference-order mode
w = bench.C3()
params, tips = w.inputs(40, 0)
bt = pgbp_b200.BatchedClusterGraphBelief(plan_of(w), 40)
bt.assignfactors(params, tips)
bt.regularizebeliefs_bycluster()
succ, iscal = bt.calibrate(None, 2)
fe = bt.factored_energy()
bt.init_beliefs_reset_fromfactors()
bt.init_messagecalibrationflags_reset()
bt.regularizebeliefs_bycluster()
succ2, _ = bt.calibrate(None, 1, reference_order=True)
assert succ.all() and succ2.all()
print("c3 ok", fe[:2, 2])

# C4 / C5 shapes at reduced size: shared-memory, multi-warp and cooperative kernels, K1 fast path p = 8 / 16
for cls, kw, B in ((bench.C4, dict(ntips=300, nretic=30), 40), (bench.C5, dict(ntips=120, nretic=12), 24)):
    w = cls(**kw)
    params, tips = w.inputs(B, 0)
    bt = pgbp_b200.BatchedClusterGraphBelief(plan_of(w), B, factors=False, residuals=False)
    bt.assignfactors(params, tips, ncolors=w.ncolors)
    succ, _ = bt.calibrate(None, 1, update_residualnorm=False)
    ll = bt.integratebelief(w.d["root_cluster"] + 1, want_mu=False)[1]
    assert succ.all() and np.isfinite(ll).all()
    print(w.key, "ok", ll[:2])
