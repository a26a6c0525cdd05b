"""CPU, world_size 2, gloo: the N>1 host logic — contiguous sharding and the
16-byte (correct, total) all-reduce.  The kernels themselves are covered by the
-m gpu tests; here the per-shard argmax comes from the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from orthogonal_classifiers_b200.distributed import (accuracy_from_counts, all_reduce_counts,
                                                          gather_predictions, shard_bounds)
    rng = np.random.default_rng(0)
    pred = rng.integers(0, 16, n)
    lab = rng.integers(0, 10, n)
    lo, hi = shard_bounds(n, rank, world)
    counts = torch.tensor([int((pred[lo:hi] == lab[lo:hi]).sum()), hi - lo], dtype=torch.int64)
    all_reduce_counts(counts)
    acc = accuracy_from_counts(counts)
    slab = torch.from_numpy(pred[lo:hi].astype(np.int64))[:, None]
    full = gather_predictions(slab, n)
    if rank == 0:
        np.save(os.path.join(tmp, "out.npy"), np.array([acc, float(counts[1]), float((pred == lab).mean()),
                                                        float(np.array_equal(full[:, 0].numpy(), pred))]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 2, 1])
def test_two_rank_count_all_reduce(tmp_path, n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    acc, total, ref, same = np.load(tmp_path / "out.npy")
    assert total == n and acc == ref and same == 1.0


def test_shard_bounds_partition():
    from orthogonal_classifiers_b200.distributed import shard_bounds
    for n in (0, 1, 7, 70_000, 1_000_000):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
