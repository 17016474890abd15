"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: contiguous shards cover the batch,
rho is broadcast once, and the sharded result assembled from both ranks equals the single-process
result (computed with the oracle here: no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from dilithium_b200.sharding import broadcast_rho, max_over_ranks, shard_range


def test_shard_range_covers_exactly():
    for total in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 8, 3) == (3 * 8192, 4 * 8192)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    import oracle_lib as ol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rho = torch.zeros(32, dtype=torch.uint8)
        if rank == 0:
            rho = torch.arange(32, dtype=torch.uint8) * 7 + 3
        broadcast_rho(rho, src=0)                       # the only collective on the path
        orc = ol.load()
        k, l = 4, 4
        a_hat = orc.expand_a(rho.numpy(), k, l)         # every rank expands A itself
        y = np.random.default_rng(99).integers(0, ol.Q, size=(total, l, 256)).astype(np.int32)
        lo, hi = shard_range(total, world, rank)
        w, _ = orc.signcore(a_hat, y[lo:hi], k, l)
        np.save(os.path.join(out_dir, f"w_{rank}.npy"), w)
        np.save(os.path.join(out_dir, f"rho_{rank}.npy"), rho.numpy())
        t = max_over_ranks(float(rank + 1))
        assert t == float(world)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_signcore_matches_single(tmp_path):
    import oracle_lib as ol
    total, world = 37, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    rho0, rho1 = np.load(tmp_path / "rho_0.npy"), np.load(tmp_path / "rho_1.npy")
    assert np.array_equal(rho0, rho1) and rho0[1] == 10
    orc = ol.load()
    a_hat = orc.expand_a(rho0, 4, 4)
    y = np.random.default_rng(99).integers(0, ol.Q, size=(total, 4, 256)).astype(np.int32)
    w_single, _ = orc.signcore(a_hat, y, 4, 4)
    w_sharded = np.concatenate([np.load(tmp_path / "w_0.npy"), np.load(tmp_path / "w_1.npy")])
    assert np.array_equal(w_single, w_sharded)
