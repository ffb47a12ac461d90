#!/usr/bin/env python
"""Benchmark of the batched Gaussian-BP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Workload (BASELINE.json configs[1]): lazaridis_2014 admixture graph, clique
tree (17 clusters / 16 sepsets / 32 messages per calibration), full
multivariate Brownian motion with p = 3 traits, fixed root, 65,536 synthetic
trait replicates per GPU simulated down the network (seed 0xB200 + 2).
Metric: clique-tree calibrations per second, whole job (all ranks).

One step =  beliefs <- factors (init_beliefs_reset_fromfactors!),
            calibrate!(beliefs, [spt]) = 16 postorder + 16 preorder messages with
            residual tracking and the iscalibrated reduction,
            integratebelief! at the root cluster (per-replicate log-likelihood),
            [N > 1: NCCL all-gather of the log-likelihoods]
for all B replicates of the rank.  `value` times that with the factors
resident in HBM; `e2e` times the public host-buffer API per step:
pinned host tip data -> H2D -> assignfactors (K1) -> calibrate -> integratebelief
-> D2H of the log-likelihoods.

`--impl reference`: the reference is Julia (not in this image), so the
reference arm is the C/OpenMP restatement of the reference's algorithm
(oracle/c, kind "port") on all host cores, on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clique-tree calibrations/sec (batched replicates)"
SEED = 0xB200


class C2:
    """BASELINE configs[1]: lazaridis_2014, MvFullBM p=3, 65,536 trait replicates, clique tree."""
    key = "c2"
    unit = "calibrations/s"
    workload = ("lazaridis_2014 admixture graph, MvFullBrownianMotion p=3, 65,536 synthetic trait replicates per GPU, "
                "clique tree, calibrate! (post+pre order, residual tracking) + integratebelief! [BASELINE configs[1]]")
    default_batch = 65536
    ncolors = 1
    residuals = True
    step_text = "reset_from_factors + calibrate (post+pre, residuals, iscal) + integrate(root)"
    e2e_text = "pinned host tip data -> assignfactors (H2D + K1) -> calibrate -> integratebelief -> D2H loglik"
    kernel_text = "k_message<i,s> family (all 32 messages of a calibration)"
    cpu_text = "assignfactors + calibrate + integratebelief per replicate"

    def __init__(self):
        self.d = json.load(open(os.path.join(ROOT, "workloads", "lazaridis_cliquetree_p3.json")))

    def inputs(self, B, rank):
        """R = A A'/3 + 0.1 I, mu = 0, fixed root; replicates simulated down the
        network: X_v = sum_k gamma_k X_pa_k + N(0, sum_k gamma_k^2 t_k R).  One theta, B data sets."""
        d = self.d
        rng = np.random.Generator(np.random.PCG64(SEED + 2 + 1000 * rank))
        p = d["ntraits"]
        A = rng.normal(size=(p, p))
        R = A @ A.T / 3 + 0.1 * np.eye(p)
        Lr = np.linalg.cholesky(R)
        n = len(d["simulate"])
        X = np.zeros((n, B, p))
        for v in range(1, n):
            par = d["simulate"][v]
            var = sum(g * g * t for _, t, g in par)
            mean = sum(g * X[q] for q, t, g in par)
            X[v] = mean + np.sqrt(var) * (rng.normal(size=(B, p)) @ Lr.T)
        tips = np.ascontiguousarray(X[d["tip_nodes"]].transpose(1, 0, 2))  # [B][ntips][p]
        params = np.concatenate([R.T.ravel(), np.zeros(p), np.zeros(p * p)])[None]
        return params, tips

    def cost(self, plan):
        by = plan.traversal_cost(0, 0, True)[0] + plan.traversal_cost(0, 1, True)[0]
        fl = plan.traversal_cost(0, 0, True)[1] + plan.traversal_cost(0, 1, True)[1]
        return by, fl

    cpu_kw = dict(post=True, pre=True, residnorm=True)


class C2S(C2):
    """C2 with the shared-precision path: the 65,536 elements are trait replicates under ONE
    parameter vector, so every J is stored and updated once; per message and element only h and g move.
    Reported separately, with its own algorithmic-bytes formula (never against the full-J byte count)."""
    key = "c2s"
    shared = True
    workload = ("lazaridis_2014 admixture graph, MvFullBrownianMotion p=3, 65,536 synthetic trait replicates per GPU under "
                "ONE parameter vector, clique tree, SHARED-PRECISION batch (J stored once per group, h and g per element), "
                "calibrate! (post+pre order, residual tracking) + integratebelief! [BASELINE configs[1], shared-J variant "
                "of SURVEY 8f-1: own byte count 8*(m_F + 1 + 4*(s+1) + s) per message]")
    kernel_text = "k_message<i,s> family, shared-precision mode (32 messages of a calibration)"

    def cost(self, plan):
        d = self.d
        dims = d["belief_dim"]
        by = 0.0
        for dr in (0, 1):
            lv = plan.levels(0, dr)
            for f, sp in zip(lv["frm"], lv["sepset"]):
                mF, s_ = dims[f], dims[sp]
                by += 8.0 * (mF + 1 + 4 * (s_ + 1) + s_)
        fl = plan.traversal_cost(0, 0, True)[1] + plan.traversal_cost(0, 1, True)[1]
        return by, fl


class C4(C2):
    """BASELINE configs[3]: synthetic 10k-tip level-1 network, HeterogeneousBM p=8, grid of 4,096 theta."""
    key = "c4"
    unit = "likelihood evaluations/s"
    workload = ("synthetic level-1 network, 10,000 tips + 1,000 reticulations (21,999 nodes; clique tree of 20,998 "
                "clusters), HeterogeneousBrownianMotion p=8 with 4 rate colours, one shared data set, batched grid of "
                "parameter vectors; unit = assignfactors! + postorder traversal + integratebelief!(root) "
                "(the optimiser objective, src/calibration.jl:195-221) [BASELINE configs[3]]")
    default_batch = 4096
    ncolors = 4
    residuals = False
    step_text = "assign_factors (K1) + propagate_1traversal_postorder (20,997 messages) + integrate(root)"
    e2e_text = "pinned host theta grid -> assignfactors (H2D + K1) -> postorder -> integratebelief -> D2H loglik"
    kernel_text = "k_message* family (20,997 postorder messages of one likelihood evaluation)"
    cpu_text = "assignfactors + postorder + integratebelief per parameter vector"

    def __init__(self, ntips=10000, nretic=1000, p=8):
        from workloads import synth
        net = synth.level1_network(ntips, nretic, SEED + 4)
        self.col = synth.edge_colors(net, self.ncolors)
        self.d = synth.cliquetree_plan(net, p, True, self.col, name="synthetic_level1_%d" % ntips)
        self.synth = synth

    def inputs(self, B, rank):
        d = self.d
        p, nc = d["ntraits"], self.ncolors
        rng = np.random.Generator(np.random.PCG64(SEED + 4 + 1000 * rank))

        def rate():
            A = rng.normal(size=(p, p))
            return A @ A.T / p + 0.1 * np.eye(p)
        R0 = [rate() for _ in range(nc)]
        # families list parents by decreasing index; simulate lists them in edge order: map through parent id
        par_col = {}
        for v, fam in enumerate(d["simulate"]):
            o = d["families"]["mem_off"][v]
            order = sorted(range(len(fam)), key=lambda k: -fam[k][0])
            for pos, k in enumerate(order):
                par_col[(v, k)] = d["families"]["mem_color"][o + 1 + pos]
        tips = self.synth.simulate_tips(d, lambda v, k: R0[par_col[(v, k)]], 1, SEED + 40)  # one shared data set
        params = np.empty((B, nc * p * p + p + p * p))
        for b in range(B):  # R_c^(b) = A A'/p + 0.1 I
            A = rng.normal(size=(nc, p, p))
            R = A @ A.transpose(0, 2, 1) / p + 0.1 * np.eye(p)
            params[b, :nc * p * p] = R.reshape(-1)  # symmetric: row- and column-major coincide
            params[b, nc * p * p:] = 0.0
        return params, tips

    def cost(self, plan):
        by, fl = plan.traversal_cost(0, 0, False)
        return by, fl

    cpu_kw = dict(post=True, pre=False, residnorm=False)


class C3(C2):
    """BASELINE configs[2]: muller_2022 network, loopy BP on the Bethe cluster graph with regularisation,
    16,384 replicates, fixed iteration count."""
    key = "c3"
    unit = "calibrations/s"
    niter = 10
    workload = ("muller_2022 network (801 nodes, 40 tips, 361 hybrids), Bethe cluster graph (1557 clusters / 1914 sepsets), "
                "UnivariateBrownianMotion(1, 0) fixed root, 16,384 simulated trait replicates per GPU, "
                "regularizebeliefs_bycluster!, spanningtrees_clusterlist schedule (2 trees), niter = 10 fixed, auto = false; "
                "1 calibration = 1 iteration over all spanning trees (4 x 1556 messages) [BASELINE configs[2]]")
    default_batch = 16384
    ncolors = 1
    residuals = True
    step_text = ("reset_from_factors + regularizebeliefs_bycluster! + calibrate!(niter = 10: 62,240 messages, residuals, "
                 "iscal) + factored_energy; value counts 10 calibrations per element per step")
    e2e_text = ("pinned host tip data -> assignfactors (H2D + K1) -> regularize -> calibrate(niter=10) -> factored_energy "
                "-> D2H")
    kernel_text = "k_message<i,s> family (62,240 messages of the 10 iterations)"
    cpu_text = "assignfactors + regularize + calibrate(niter=10) + factored_energy per replicate"

    def __init__(self):
        self.d = json.load(open(os.path.join(ROOT, "workloads", "muller_bethe_p1.json")))

    def inputs(self, B, rank):
        d = self.d
        rng = np.random.Generator(np.random.PCG64(SEED + 3 + 1000 * rank))
        n = len(d["simulate"])
        X = np.zeros((n, B, 1))
        for v in range(1, n):
            par = d["simulate"][v]
            var = sum(g * g * t for _, t, g in par)
            mean = sum(g * X[q] for q, t, g in par)
            X[v] = mean + np.sqrt(var) * rng.normal(size=(B, 1))
        tips = np.ascontiguousarray(X[d["tip_nodes"]].transpose(1, 0, 2))
        params = np.array([[1.0, 0.0, 0.0]])
        return params, tips

    def cost(self, plan):
        nt = len(self.d["trees"])
        by = sum(plan.traversal_cost(t, dr, True)[0] for t in range(nt) for dr in range(2)) * self.niter
        fl = sum(plan.traversal_cost(t, dr, True)[1] for t in range(nt) for dr in range(2)) * self.niter
        return by, fl

    cpu_kw = dict(post=True, pre=True, residnorm=True, niter=10, reg_bycluster=True)


class C3L(C3):
    """BASELINE configs[2], second cluster graph: LTRIP(net) (the node families as clusters: 801 clusters / 1158
    sepsets, sepsets of up to two nodes)."""
    key = "c3l"
    workload = ("muller_2022 network (801 nodes, 40 tips, 361 hybrids), LTRIP(net) cluster graph (801 clusters / 1158 sepsets), "
                "UnivariateBrownianMotion(1, 0) fixed root, 16,384 simulated trait replicates per GPU, "
                "regularizebeliefs_bycluster!, spanningtrees_clusterlist schedule (2 trees), niter = 10 fixed, auto = false; "
                "1 calibration = 1 iteration over all spanning trees (4 x 800 messages) [BASELINE configs[2], LTRIP variant]")
    step_text = ("reset_from_factors + regularizebeliefs_bycluster! + calibrate!(niter = 10: 32,000 messages, residuals, "
                 "iscal) + factored_energy; value counts 10 calibrations per element per step")
    kernel_text = "k_tilewalk / k_message<i,s> family (32,000 messages of the 10 iterations)"

    def __init__(self):
        self.d = json.load(open(os.path.join(ROOT, "workloads", "muller_ltrip_p1.json")))


class C5(C4):
    """BASELINE configs[4]: synthetic 100k-tip network with 10k reticulations, MvFullBrownianMotion p=16, one
    theta, replicate batch as large as HBM allows per GPU (0.92 GB of state per replicate), sharded by
    replicate across GPUs."""
    key = "c5"
    unit = "calibrations/s"
    default_batch = 128
    ncolors = 1
    residuals = False
    workload = ("synthetic level-1 network, 100,000 tips + 10,000 reticulations (219,999 nodes; clique tree of 209,998 "
                "clusters, sender dimensions 16 / 32 / 48), MvFullBrownianMotion p=16, one parameter vector, 128 simulated "
                "trait replicates per GPU (0.92 GB of state each: 118 GB of the 180 GB), assignfactors! + calibrate! "
                "(post+pre order) + integratebelief!(root) [BASELINE configs[4]]")
    step_text = "assign_factors (K1, generic path: p = 16) + calibrate (419,994 messages) + integrate(root)"
    e2e_text = "pinned host tip data -> assignfactors (H2D + K1) -> calibrate -> integratebelief -> D2H loglik"
    kernel_text = "k_message* family (419,994 messages of one calibration; (16,16) shared-memory kernel, (32,16) generic)"
    cpu_text = "assignfactors + calibrate + integratebelief per replicate"

    def __init__(self, ntips=100000, nretic=10000, p=16):
        from workloads import synth
        net = synth.level1_network(ntips, nretic, SEED + 5)
        self.col = None
        self.d = synth.cliquetree_plan(net, p, True, None, name="synthetic_level1_%d" % ntips)
        self.synth = synth

    def inputs(self, B, rank):
        d = self.d
        p = d["ntraits"]
        rng = np.random.Generator(np.random.PCG64(SEED + 5 + 1000 * rank))
        A = rng.normal(size=(p, p))
        R = A @ A.T / p + 0.1 * np.eye(p)
        tips = self.synth.simulate_tips(d, lambda v, k: R, B, SEED + 50 + rank)
        params = np.concatenate([R.reshape(-1), np.zeros(p), np.zeros(p * p)])[None]
        return params, tips

    def cost(self, plan):
        by = plan.traversal_cost(0, 0, False)[0] + plan.traversal_cost(0, 1, False)[0]
        fl = plan.traversal_cost(0, 0, False)[1] + plan.traversal_cost(0, 1, False)[1]
        return by, fl

    cpu_kw = dict(post=True, pre=True, residnorm=False)


WORKLOADS = {"c2": C2, "c2s": C2S, "c3": C3, "c3l": C3L, "c4": C4, "c5": C5}


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_port_rate(w, params, tips, seconds, nthreads=0, steps=1, warmup=0):
    """units/s of the C/OpenMP oracle port on a bounded sample of the workload."""
    from oracle.cport import COracle, dll
    co = COracle.from_plan_dict(w.d)
    kw = dict(root_belief=w.d["root_cluster"], nthreads=nthreads, ncolors=w.ncolors, want_fe=(w.key in ("c3", "c3l")), **w.cpu_kw)
    B = max(params.shape[0], tips.shape[0])

    def run(n):
        return co.run_batch(params[:n] if params.shape[0] > 1 else params, tips[:n] if tips.shape[0] > 1 else tips,
                            B=n, **kw)
    n0 = min(max(64, B // 32), B)
    run(n0)  # warm-up (thread pool, page faults)
    t = time.perf_counter()
    run(n0)
    r0 = n0 / (time.perf_counter() - t)
    n = int(min(B, max(n0, r0 * seconds)))
    for _ in range(warmup):
        run(n)
    t = time.perf_counter()
    for _ in range(steps):
        out = run(n)
    dt = (time.perf_counter() - t) / steps
    assert (out["status"] == 0).all()
    cores = dll().pgbpo_num_threads() if nthreads <= 0 else nthreads
    return n / dt, cores, n, dt, (out["fe"][:, 2] if w.key in ("c3", "c3l") else out["loglik"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]()
    B = args.batch or w.default_batch
    params, tips = w.inputs(B, 0)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it explicitly)
    nthr = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    value, cores, n, dt, _ = cpu_port_rate(w, params, tips, 3.0, nthreads=nthr, steps=args.steps, warmup=args.warmup)
    value *= getattr(w, "niter", 1)
    sample = f"{n} of {B} batch elements per step ({w.cpu_text})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": w.unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.workload, "note": "reference is Julia (absent from the image): C/OpenMP restatement of "
                   "its algorithm (oracle/c), all host threads, bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": w.unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": w.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled in-process every
    2 ms (the timed region of the small config lasts ~20 ms, shorter than nvidia-smi's period);
    falls back to `nvidia-smi -lms 100` if NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.nv, self.stop_flag, self.mx = [], None, None, False, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                c = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), c, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                c, m = float(f[0]), float(f[1])
            except (ValueError, IndexError):
                continue
            self.mx = m
            self.rows.append((time.perf_counter(), c, [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))

    def stop(self, t0, t1):
        if self.nv is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        self.stop_flag = True
        inside = [(c, r) for t, c, r in self.rows if t0 <= t <= t1]
        where = "timed region"
        if not inside:  # only possible with the nvidia-smi fallback on a very short region
            inside, where = [(c, r) for t, c, r in self.rows], "whole run (region shorter than the sampling period)"
        sm = [c for c, _ in inside]
        reasons = sorted({x for _, r in inside for x in r})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "sampled": where, "source": "nvml" if self.nv else "nvidia-smi"}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import pgbp_b200
    from pgbp_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={world}", file=sys.stderr)
    w = WORKLOADS[args.workload]()
    d = w.d
    B = args.batch or w.default_batch
    p = d["ntraits"]
    params, tips = w.inputs(B, rank)
    lib = pgbp_b200.default_library()
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"], p,
                                      d["families"], lib)
    stream = torch.cuda.current_stream()
    group = B if getattr(w, "shared", False) else 0
    bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, device=local, stream=stream.cuda_stream,
                                             factors=w.residuals, residuals=w.residuals, shared_precision_group=group)
    if args.walk is not None:
        bt.set_walk_mode(args.walk)
    if args.pipeline is not None:
        bt.set_pipeline(args.pipeline)
    if args.graph is not None:
        bt.set_graph_mode(args.graph)
    if args.tilewalk is not None:
        bt.set_tilewalk_mode(args.tilewalk)
    if args.coop is not None:
        bt.set_coop_mode(args.coop)
    if args.tw_lanes or args.tw_wide:
        bt.set_tilewalk_params(args.tw_lanes, args.tw_wide)
    root = d["root_cluster"] + 1
    bytes_unit, flops_unit = w.cost(plan)
    nmsg = ({"c2": 2 * len(d["trees"][0][0]), "c2s": 2 * len(d["trees"][0][0]), "c4": len(d["trees"][0][0]),
             "c5": 2 * len(d["trees"][0][0])}.get(w.key)
            or 2 * sum(len(t[0]) for t in d["trees"]) * w.niter)
    upe = getattr(w, "niter", 1)  # metric units per element per step

    # ---- device-resident arm: inputs in HBM before the timed region -----------------------
    _, ld, _ = bt.device_view()
    # two result buffers: the NCCL gather of step k (on NCCL's stream) overlaps the kernels of step k+1
    d_norms = [torch.empty(ld, dtype=torch.float64, device=dev) for _ in range(2)]
    gathered = [torch.empty(world * ld, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    pending = [None, None]
    counter = [0]

    def finish(ev=None):
        """integratebelief! at the root + (N > 1) asynchronous all-gather of the log-likelihoods."""
        k = counter[0] % 2
        counter[0] += 1
        if pending[k] is not None:
            pending[k].wait()  # stream-level wait: buffer k is free again
            pending[k] = None
        bt.integrate_device(root, d_norms[k].data_ptr())
        if world > 1:
            pending[k] = dist.all_gather_into_tensor(gathered[k], d_norms[k], async_op=True)
        return d_norms[k]
    if w.key in ("c2", "c2s"):
        bt.assignfactors(params, tips)  # factors resident in HBM

        def step(ev=None):
            bt.init_beliefs_reset_fromfactors()
            if ev:
                ev[0].record(stream)
            bt.calibrate_async(None, 1, update_residualnorm=True)
            if ev:
                ev[1].record(stream)
            finish()
    elif w.key in ("c3", "c3l"):
        bt.assignfactors(params, tips)
        d_fe = torch.empty(3 * ld, dtype=torch.float64, device=dev)

        def finish(ev=None):  # noqa: F811  (loopy objective: factored energy instead of integratebelief!)
            k = counter[0] % 2
            counter[0] += 1
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
            bt.factored_energy_device(d_fe.data_ptr())
            d_norms[k].copy_(d_fe[2 * ld:3 * ld])
            if world > 1:
                pending[k] = dist.all_gather_into_tensor(gathered[k], d_norms[k], async_op=True)

        def step(ev=None):
            bt.init_beliefs_reset_fromfactors()
            bt.init_messagecalibrationflags_reset()
            bt.regularizebeliefs_bycluster()
            if ev:
                ev[0].record(stream)
            bt.calibrate_async(None, w.niter, update_residualnorm=True)
            if ev:
                ev[1].record(stream)
            finish()
    else:
        d_params = torch.from_numpy(params).to(dev)
        d_tips = torch.from_numpy(tips).to(dev)
        direction = L.CAL_BOTH if w.key == "c5" else L.CAL_POSTORDER

        def step(ev=None):
            bt.assignfactors_device(d_params.data_ptr(), params.shape[0], d_tips.data_ptr(), tips.shape[0], ncolors=w.ncolors)
            if ev:
                ev[0].record(stream)
            bt.calibrate_async(None, 1, update_residualnorm=False, direction=direction)
            if ev:
                ev[1].record(stream)
            finish()

    def barrier():
        for k in range(2):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    bt.launch_count(reset=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(args.steps):
        step(evs[k])
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    launches = bt.launch_count()
    ms_total = e0.elapsed_time(e1)
    ms_msgs = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop(t0, t1) if sampler else None
    tt = torch.tensor([ms_total, ms_msgs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, ms_msgs = float(tt[0]), float(tt[1])
    value = world * B * upe * args.steps / (ms_total * 1e-3)
    st = bt.status()
    assert (st == 0).all(), "numerical failure inside the timed region"
    loglik_dev = d_norms[(counter[0] - 1) % 2][:B].cpu().numpy()
    if world > 1:  # the gathered vector holds every rank's log-likelihoods; this rank's slice must match
        mine = gathered[(counter[0] - 1) % 2].view(world, ld)[rank, :B].cpu().numpy()
        assert np.array_equal(mine, loglik_dev)

    # ---- end-to-end arm: public host-buffer API, pinned host inputs, D2H of the result -------
    # Every step = one synchronous sequence of public calls on one batch: H2D of that step's inputs
    # (pinned) + K1, message passing, integratebelief! with the D2H of the result.  When two batches
    # fit in HBM the steps alternate between two batches driven by two host threads (the ABI allows
    # distinct batches on distinct threads): the H2D / D2H of one step overlaps the kernels of the other.
    big = params if w.key == "c4" else tips
    e2e_steps = args.steps if w.key in ("c2", "c2s") else min(args.steps, 5)
    free_b, total_b = torch.cuda.mem_get_info()
    nb_e2e = max(1, min(args.e2e_batches, 1 + int(free_b / (bt.device_bytes() * 1.1))))
    two = nb_e2e > 1
    bts = [bt]
    for _ in range(nb_e2e - 1):
        bts.append(pgbp_b200.BatchedClusterGraphBelief(plan, B, device=local, factors=w.residuals, residuals=w.residuals,
                                                       shared_precision_group=group))
        if args.pipeline is not None:
            bts[-1].set_pipeline(args.pipeline)
    pin_np = [torch.from_numpy(big.copy()).pin_memory().numpy() for _ in bts]
    results = [None] * len(bts)

    def e2e_step(i):
        b_ = bts[i]
        if w.key in ("c2", "c2s"):
            b_.assignfactors(params, pin_np[i])                 # H2D of this step's inputs + K1
            succ, iscal = b_.calibrate(None, 1)                 # D2H of succ / iscal
        elif w.key in ("c3", "c3l"):
            b_.assignfactors(params, pin_np[i])
            b_.regularizebeliefs_bycluster()
            succ, iscal = b_.calibrate(None, w.niter)
            results[i] = b_.factored_energy()[:, 2]
            return
        elif w.key == "c5":
            b_.assignfactors(params, pin_np[i])
            succ, _ = b_.calibrate(None, 1, update_residualnorm=False)
        else:
            b_.assignfactors(pin_np[i], tips, ncolors=w.ncolors)
            succ = b_.propagate_1traversal_postorder(0, update_residualnorm=False)
        results[i] = b_.integratebelief(root, want_mu=False)[1]  # D2H of the result

    def worker(i, n):
        torch.cuda.set_device(local)
        for _ in range(n):
            e2e_step(i)

    def run_e2e(nsteps):
        if len(bts) == 1:
            worker(0, nsteps)
            return
        nb = len(bts)
        ths = [threading.Thread(target=worker, args=(i, nsteps // nb + (i < nsteps % nb))) for i in range(nb)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    run_e2e(2 * len(bts))
    barrier()
    t0e = time.perf_counter()
    run_e2e(e2e_steps)
    barrier()
    dte = time.perf_counter() - t0e
    te = torch.tensor([dte], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * upe * e2e_steps / float(te[0])
    h2d = tips.nbytes + params.nbytes
    ll_host = results[0]
    d2h = ll_host.nbytes + (1 if w.key == "c4" else 2) * 4 * B
    for r_ in results:
        assert np.allclose(r_, loglik_dev, rtol=1e-12, atol=0)  # same kernels, same inputs: identical
    del bts[1:]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (k_message*) ----------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = bytes_unit * B * args.steps / (ms_msgs * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(w.key, {}).get("dram_bytes_per_unit_per_element")
        if traffic is not None:
            traffic = traffic * B
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": w.kernel_text,
                "algorithmic_bytes_per_unit_per_element": bytes_unit,
                "algorithmic_flops_per_unit_per_element": flops_unit,
                "fp64_gflops_achieved": flops_unit * B * args.steps / (ms_msgs * 1e-3) / 1e9,
                "share_of_step": ms_msgs / ms_total}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        nthr = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        rate, cores, n, dt, ll_cpu = cpu_port_rate(w, params, tips, args.cpu_seconds, nthreads=nthr)
        rate *= upe
        err = float(np.max(np.abs(ll_cpu / loglik_dev[:n] - 1)))
        cpu = {"value": rate, "unit": w.unit, "cores": cores, "kind": "port",
               "sample": f"{n} of {B} batch elements in {dt:.1f} s ({w.cpu_text}, C/OpenMP restatement of the Julia "
                         f"reference)",
               "max_rel_err_gpu_vs_cpu_loglik": err}
        # (c3: loopy BP with eps = 2.2e-16 regularisation of factor-less clusters is ill-conditioned: restatements
        # of the reference's own formulation differ by 6e-7 already; see tests/test_fullsize.py)
        assert err < (2e-4 if w.key in ("c3", "c3l") else 1e-10), err

    line = {
        "metric": METRIC, "value": value, "unit": w.unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.workload, "batch_per_gpu": B, "ntraits": p, "messages_per_unit": nmsg,
                   "step": w.step_text + (" + nccl all_gather(loglik), double-buffered and asynchronous: it overlaps "
                                          "the next step's kernels" if world > 1 else ""),
                   "l2": "inputs larger than L2 (state %.2f GB per GPU)" % (bt.device_bytes() / 1e9),
                   "parallelism": f"batch sharded over {world} GPU(s), plan replicated"},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": w.unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "path": w.e2e_text + ((" (%d batches x %d host threads: copies of one step overlap the "
                                                          "kernels of the others)" % (nb_e2e, nb_e2e)) if two else "")},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="batch elements per GPU (default: the workload's)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE configs[1] (headline), "
                    "c4 = configs[3] (10k-tip synthetic network, p=8)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=4, help="batches (= host threads) alternating in the end-to-end arm")
    ap.add_argument("--pipeline", type=int, default=None, help="element chunks of a calibration (-1 auto, 1 off)")
    ap.add_argument("--tilewalk", type=int, default=None, help="tile-walk kernel (-1 auto, 0 off, 1 on)")
    ap.add_argument("--tw-lanes", type=int, default=0, help="tile-walk message lanes per block (4, 8, 16)")
    ap.add_argument("--tw-wide", type=int, default=0, help="tile-walk: steps wider than this keep their own launches")
    ap.add_argument("--coop", type=int, default=None, help="medium-shape kernel: -1 auto, 1 shared-memory, 8 cooperative, 0 generic")
    ap.add_argument("--graph", type=int, default=None, help="CUDA-graph replay of calibrate (-1 auto, 0 off, 1 on)")
    ap.add_argument("--walk", type=int, default=None, help="kernel strategy override: 0 level-parallel, 1 walk kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
