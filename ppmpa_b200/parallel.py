"""Multi-GPU plumbing of the progressive pass loop.

PPM-PA passes share nothing but the read-only scene (the reference runs them as separate
processes, util/iterator.rb:90-117, and sums the images offline, util/averager2.rb:49-62,86).
So the path shards by pass: pass i goes to rank i mod G with its own Philox stream (seed, i)
and its own radius r_i; every rank keeps a local sum image + pass count, and ONE sum-reduce of
(3*W*H + 1) doubles per frame combines them.  No other exchange exists on this path.
"""
import numpy as np


def passes_for_rank(n_pass, world, rank):
    """Pass indices rendered by `rank`: {i : i mod world == rank}, ascending."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_pass, world))


def reduce_accumulators(acc, dst=0, group=None):
    """In-place sum-reduce of the [sum image | pass count] vector onto rank `dst`.
    `acc` is a torch tensor (CUDA with the nccl backend, CPU with gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return acc


def mean_image(acc):
    """sum / number of passes actually summed (averager2.rb:86 divides by the file count;
    a missing pass simply does not count, averager2.rb:65)."""
    a = np.asarray(acc.cpu() if hasattr(acc, "cpu") else acc, dtype=np.float64)
    n = a[-1]
    if n <= 0:
        raise ValueError("no pass accumulated")
    return a[:-1].reshape(-1, 3) / n, int(n)
