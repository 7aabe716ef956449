"""Batch sharding for N GPUs (one process per GPU).  Utterances are independent
(SURVEY.md §8(e)): every rank scores a contiguous shard of the batch with no data-path
collective; the only exchange is one all-reduce of the scalar loss (and, for ASG /
transducer-with-transitions, of the small transition gradient, which DDP already does for
the criterion's parameters in the reference, train.py:205-208)."""
import torch
import torch.distributed as dist


def shard_bounds(batch_size, rank, world_size):
    """[lo, hi) of rank's contiguous shard; sizes differ by at most one."""
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_shards(costs, world_size):
    """Greedy longest-processing-time assignment of utterances to ranks by cost (e.g.
    T * (2 L_b + 1) for ragged targets); returns a list of index lists."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: loads[k])
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def global_mean_loss(local_loss_sum, global_batch, group=None):
    """mean over the GLOBAL batch of per-utterance (scaled) losses, given this rank's sum:
    one all-reduce of a single float (ncclAllReduce SUM over NVLink; gloo on CPU)."""
    t = local_loss_sum.detach().reshape(1).clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0] / global_batch
