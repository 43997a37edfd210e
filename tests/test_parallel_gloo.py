"""world_size-2 gloo test of the multi-GPU host logic: pass sharding + the single sum-reduce
per frame.  Per-rank images are synthetic functions of the pass index."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppmpa_b200 import parallel

NPIX = 37
NPASS = 7


def fake_pass_image(i):
    rng = np.random.default_rng(1000 + i)
    return rng.random((NPIX, 3))


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc = torch.zeros(NPIX * 3 + 1, dtype=torch.float64)
    for i in parallel.passes_for_rank(NPASS, world, rank):
        acc[:-1] += torch.from_numpy(fake_pass_image(i).reshape(-1))
        acc[-1] += 1.0
    parallel.reduce_accumulators(acc, dst=0)
    if rank == 0:
        mean, n = parallel.mean_image(acc)
        np.save(out, np.concatenate([mean.reshape(-1), [n]]))
    dist.destroy_process_group()


def test_pass_sharding_is_a_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in parallel.passes_for_rank(NPASS, world, r))
        assert seen == list(range(NPASS))
    assert parallel.passes_for_rank(10, 4, 1) == [1, 5, 9]
    with pytest.raises(ValueError):
        parallel.passes_for_rank(4, 2, 2)


def test_two_rank_reduce_matches_serial_sum(tmp_path):
    out = str(tmp_path / "res.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    res = np.load(out)
    want = sum(fake_pass_image(i) for i in range(NPASS)) / NPASS
    assert int(res[-1]) == NPASS
    assert np.allclose(res[:-1].reshape(-1, 3), want, rtol=1e-15, atol=0)


def test_mean_image_skips_missing_passes():
    acc = np.concatenate([np.full(6, 6.0), [3.0]])
    mean, n = parallel.mean_image(acc)
    assert n == 3 and np.all(mean == 2.0)
    with pytest.raises(ValueError):
        parallel.mean_image(np.zeros(7))
