"""N > 1 path on the CPU: two ranks (gloo, 127.0.0.1), each owning a contiguous slice of the batch,
the plan replicated, one all-gather of the per-element results.  The ranks run the host-emulation
build of the library (this is a test of the host-side sharding / gather logic; the GPU arm is
bench.py --gpus N).  The gathered log-likelihoods must equal a single-rank run bit for bit."""
import json
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(B):
    import pgbp_b200  # noqa: F401
    from harness import Case, get_lib
    from oracle import models as M
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")))
    lib = get_lib("emul")
    p = 2
    rng = np.random.default_rng(77)
    A = rng.normal(size=(p, p))
    R = A @ A.T / p + 0.1 * np.eye(p)
    taxa = ["Mbuti", "Onge", "Karitiana", "MA1", "Loschbour", "European", "Stuttgart"]
    data = rng.normal(size=(B, 7, p))
    model = M.MvFullBrownianMotion(R, np.zeros(p))
    case = Case(gold["lazaridis"], "cliquetree", data[0], taxa, model, lib, order_hint=gold["lazaridis_cluster_labels"])
    return case, R, data


def _run(case, R, data):
    import pgbp_b200
    p = R.shape[0]
    bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, data.shape[0])
    bt.assignfactors(pgbp_b200.bm_params([R], np.zeros(p)), data)
    succ, iscal = bt.calibrate(case.sched)
    ll = bt.integratebelief(case.sched[0][2][0])[1]
    fe = bt.factored_energy()
    return np.stack([ll, fe[:, 2], succ.astype(float), iscal.astype(float)])


def _worker(rank, world, port, B, outdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pgbp_b200 import sharding
        case, R, data = _problem(B)
        (mine,) = sharding.shard_inputs([data], B, rank, world)
        assert mine.shape[0] == len(range(*sharding.shard_slice(B, rank, world).indices(B)))
        local = _run(case, R, mine) if mine.shape[0] else np.zeros((4, 0))
        full = sharding.allgather_elements(torch.from_numpy(np.ascontiguousarray(local)), B)
        np.save(os.path.join(outdir, f"rank{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,world", [(11, 2), (3, 2), (1, 2)])
def test_two_ranks_gloo_match_single_rank(tmp_path, B, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    case, R, data = _problem(B)
    ref = _run(case, R, data)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == ref.shape
        assert np.array_equal(got, ref)  # same kernels, same inputs per element: bit-identical
    assert ref[2].all()  # succ; (iscal needs a second pass: residuals of the first one are not small)


def test_shard_slices_cover_batch():
    from pgbp_b200 import sharding
    for B in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                sl = sharding.shard_slice(B, r, world)
                seen.extend(range(sl.start, sl.stop))
                assert sl.stop - sl.start <= sharding.shard_size(B, world)
            assert seen == list(range(B))


def test_peer_gather_two_ranks_in_process():
    # The fused integratebelief! + gather (pgbp_comm_*): two "ranks" with their own batch and window in one
    # process on the host-emulation build (handles are plain pointers there); every window must end up holding
    # both ranks' log-likelihoods in rank order, equal to the ordinary integratebelief! results; a wait before the
    # peer has put is an error.
    import pgbp_b200
    from harness import get_lib
    from pgbp_b200 import sharding
    B, world = 11, 2
    case, R, data = _problem(B)
    lib = get_lib("emul")
    p = R.shape[0]
    shards = [sharding.shard_inputs([data], B, r, world)[0] for r in range(world)]
    ld = 32
    comms = [sharding.PeerGather(lib, 0, r, world, ld, nbuffers=2, connect=False) for r in range(world)]
    handles = [c.ipc_handle for c in comms]
    for c in comms:
        c.connect(handles)
    bts, ref = [], []
    for r in range(world):
        bt = pgbp_b200.BatchedClusterGraphBelief(case.plan, shards[r].shape[0])
        bt.assignfactors(pgbp_b200.bm_params([R], np.zeros(p)), shards[r])
        bt.calibrate(case.sched)
        ref.append(bt.integratebelief(case.sched[0][2][0])[1])
        bts.append(bt)
    comms[0].integrate_gather(bts[0], case.sched[0][2][0], 1)
    with pytest.raises(pgbp_b200.PgbpError):
        comms[0].wait(bts[0], 1)  # rank 1 has not put buffer 1 yet
    comms[1].integrate_gather(bts[1], case.sched[0][2][0], 1)
    for r in range(world):
        comms[r].wait(bts[r], 1)
        comms[r].check(bts[r])
        win = comms[r].read(bts[r], 1)
        for q in range(world):
            assert np.array_equal(win[q, :shards[q].shape[0]], ref[q])
        assert np.array_equal(comms[r].read(bts[r], 0), np.zeros((world, ld)))
    # generic put of a device vector (here: the same log-likelihoods) into buffer 0
    import ctypes as C
    for r in range(world):
        src = np.ascontiguousarray(ref[r] * 2.0)
        comms[r].put(bts[r], 0, src.ctypes.data_as(C.c_void_p).value)
    for r in range(world):
        comms[r].wait(bts[r], 0)
        win = comms[r].read(bts[r], 0)
        for q in range(world):
            assert np.array_equal(win[q, :shards[q].shape[0]], 2.0 * ref[q])
    with pytest.raises(pgbp_b200.PgbpError):
        sharding.PeerGather(lib, 0, 0, 1, 4, connect=False).integrate_gather(bts[0], 1, 0)  # window shorter than the batch
    for c in comms:
        c.close()


@pytest.mark.gpu
def test_two_devices_in_one_process_raise_their_own_kernel_limits():
    # The kernels that need more than 48 KB of dynamic shared memory (k_message_smem*, k_hmsg_bulk, k_jwalk) get their
    # limit raised per DEVICE (cudaFuncSetAttribute applies to the current device only): a process that drives two
    # GPUs, as the Julia wrapper may, must be able to run the same shapes on the second one.  p = 16 on a small
    # level-1 network: ordinary batch (shared-memory kernels) and shared-precision batch (bulk-copy element pass) on
    # device 0, then on device 1; identical results.
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import bench
    from harness import get_lib
    import pgbp_b200
    lib = get_lib("cuda")
    w = bench.C5(ntips=60, nretic=6)
    d = w.d
    B = 200
    params, tips = w.inputs(B, 0)
    plan = pgbp_b200.ClusterGraphPlan(d["nclusters"], d["belief_dim"], d["sepset_clusters"], d["upind"], d["trees"],
                                      d["ntraits"], d["families"], lib)
    root = d["root_cluster"] + 1
    out = {}
    for dev in (0, 1):
        for group in (0, B):
            bt = pgbp_b200.BatchedClusterGraphBelief(plan, B, device=dev, shared_precision_group=group)
            bt.assignfactors(params, tips)
            succ, _ = bt.calibrate(None, 1)
            assert succ.all()
            out[dev, group] = bt.integratebelief(root)[1]
    assert np.array_equal(out[0, 0], out[1, 0]) and np.array_equal(out[0, B], out[1, B])
    assert np.array_equal(out[0, 0], out[0, B])
