import sys, time, json, threading
sys.path.insert(0, '/root/repo')
import numpy as np, torch, bench, pgbp_b200
w = bench.C2(); d = w.d; B = 65536
params, tips = w.inputs(B, 0)
lib = pgbp_b200.default_library()
plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"], 3, d["families"], lib)
root = d["root_cluster"] + 1
pin = torch.from_numpy(tips.copy()).pin_memory().numpy()
def mk():
    return pgbp_b200.BatchedClusterGraphBelief(plan, B)
bt = mk()
def step(b_, buf):
    t0 = time.perf_counter(); b_.assignfactors(params, buf)
    t1 = time.perf_counter(); b_.calibrate(None, 1)
    t2 = time.perf_counter(); r = b_.integratebelief(root, want_mu=False)[1]
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for _ in range(5): step(bt, pin)
acc = np.zeros(3)
for _ in range(40): acc += step(bt, pin)
print("single thread ms: assign %.3f calibrate %.3f integrate %.3f total %.3f" % (*(acc / 40 * 1e3), acc.sum() / 40 * 1e3))
for nth in (2, 3, 4):
    bts = [mk() for _ in range(nth)]
    pins = [torch.from_numpy(tips.copy()).pin_memory().numpy() for _ in range(nth)]
    def work(i, n):
        for _ in range(n): step(bts[i], pins[i])
    for i in range(nth): work(i, 3)
    torch.cuda.synchronize()
    t = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i, 40)) for i in range(nth)]
    [x.start() for x in ths]; [x.join() for x in ths]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(nth, "threads: %.1f M calibrations/s, %.3f ms/step" % (nth * 40 * B / dt / 1e6, dt / (nth * 40) * 1e3))
    del bts
