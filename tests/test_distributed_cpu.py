"""CPU test of the N>1 host logic (gloo, world_size 2): contiguous batch shards plus one
all-reduce of the scalar loss reproduce the single-process batch mean; the per-shard
gradient scale uses the GLOBAL batch size.  The per-utterance numbers come from the
oracle (there is no GPU here); the sharding / reduction code is the product's."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["GTN_ORACLE_THREADS"] = "1"
    import gtn64
    import ref_criterions as rc
    from gtn_applications_b200.distributed import shard_bounds, global_mean_loss
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    B, T, C = 7, 12, 5
    x = rng.standard_normal((B, T, C)).astype(np.float32)
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in (3, 0, 5, 2, 4, 1, 3)]
    lo, hi = shard_bounds(B, rank, world)
    res = rc.ctc(gtn64, x[lo:hi], tg[lo:hi], C - 1, "mean")
    local_sum = torch.tensor(res["losses"].sum())
    loss = global_mean_loss(local_sum, B)
    grad = res["grad"] * (hi - lo) / B        # oracle divides by the shard size; rescale to global B
    out[rank] = (loss.item(), lo, hi, grad)
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_single_process_mean():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gtn64
    import ref_criterions as rc
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    rng = np.random.default_rng(0)
    B, T, C = 7, 12, 5
    x = rng.standard_normal((B, T, C)).astype(np.float32)
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in (3, 0, 5, 2, 4, 1, 3)]
    full = rc.ctc(gtn64, x, tg, C - 1, "mean")
    assert abs(out[0][0] - full["loss"]) < 1e-12 and abs(out[1][0] - full["loss"]) < 1e-12
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 4, 4, 7)
    np.testing.assert_allclose(np.concatenate([out[0][3], out[1][3]]), full["grad"], rtol=1e-12)


def test_shard_helpers():
    from gtn_applications_b200.distributed import shard_bounds, balanced_shards
    for B in (1, 7, 256, 2048):
        for W in (1, 2, 4, 8):
            b = [shard_bounds(B, r, W) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    shards = balanced_shards([10, 1, 1, 1, 9, 2, 2, 6], 2)
    assert sorted(i for s in shards for i in s) == list(range(8))
    loads = [sum([10, 1, 1, 1, 9, 2, 2, 6][i] for i in s) for s in shards]
    assert abs(loads[0] - loads[1]) <= 2
