"""N > 1 path on CPU: chain sharding and the final all-reduce over torch.distributed (gloo, world_size 2).
The sweep itself needs a GPU, so the ranks exchange synthetic per-chain results; what is covered is the
host logic of `ParallelProcessManager` that runs around the kernels: contiguous block sharding, global chain
offsets for the RNG stream, and the SUM all-reduce that must equal a host `np.sum`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latticeqmc_b200.multiprocessing import shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 256, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    rs = [np.random.RandomState(1000 + c) for c in range(lo, hi)]           # keyed by GLOBAL chain index
    means = np.zeros((total, 2, n, n))
    for k, c in enumerate(range(lo, hi)):
        means[c] = rs[k].rand(2, n, n)
    buf = torch.from_numpy(means.ravel().copy())
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, buf.numpy().reshape(means.shape))
    dist.destroy_process_group()


def test_allreduce_equals_concatenation(tmp_path):
    """A 2-rank run equals the concatenation of the per-chain results of a 1-rank run, and the manager's
    unweighted mean over chains (multiprocessing.py:265-267) equals the host np.sum / procs."""
    total, n, world = 5, 3, 2
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, n, out), nprocs=world, join=True)
    got = np.load(out)
    ref = np.stack([np.random.RandomState(1000 + c).rand(2, n, n) for c in range(total)])
    assert np.array_equal(got, ref)
    assert np.allclose(got.sum(0) / total, np.sum(ref, axis=0) / total, rtol=0, atol=0)


def test_manager_job_split_matches_reference():
    """set_jobs: sweeps/procs each, remainder to chain 0 (multiprocessing.py:260-263); bad job lists raise
    ValueError (multiprocessing.py:107-108)."""
    from latticeqmc_b200 import HubbardModel, ParallelProcessManager
    from latticeqmc_b200.multiprocessing import ProcessManager
    model = HubbardModel(u=4, t=1)
    model.build_square(2)
    mgr = ParallelProcessManager(model, 2.0, 20, warmup=10, procs=3, seeds=[1, 2, 3])
    mgr.set_jobs(100)
    assert list(mgr.var_kwargs["sweeps"]) == [34, 33, 33] and mgr.total == 3
    pm = ProcessManager(procs=2)
    with pytest.raises(ValueError):
        pm.set_jobs(a=[1, 2], b=[1])
