"""N>1 host logic on CPU: world_size-2 gloo run of the realisation sharding and the
single histogram all-reduce (the only collective of the hot path)."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from wavelet_transformer_b200 import engine

DT = 1 / 12


def test_shard_range_partitions():
    for total in (0, 1, 7, 300, 100000):
        for world in (1, 2, 3, 4, 8):
            parts = [engine.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        engine.shard_range(10, 2, 2)


def _fake_hist(first, count, S=21):
    """Deterministic per-realisation 'histogram' keyed by the GLOBAL index."""
    h = np.zeros((S, 1000), dtype=np.int64)
    for m in range(first, first + count):
        rng = np.random.default_rng(m)
        idx = rng.integers(300, 1000, (S, 200))
        for s in range(S):
            np.add.at(h[s], idx[s], 1)
    return h


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sig, hist = engine.wct_significance_sharded(0.8, 0.6, DT, 1 / 4, 2 * DT, 20, mc_count=11, hist_fn=_fake_hist)
        out[rank] = (sig, hist)
    finally:
        dist.destroy_process_group()


def test_sharded_significance_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    single_sig, single_hist = engine.wct_significance_sharded(0.8, 0.6, DT, 1 / 4, 2 * DT, 20, mc_count=11,
                                                              hist_fn=_fake_hist)
    assert np.array_equal(single_hist.astype(np.int64), _fake_hist(0, 11))
    for rank in (0, 1):
        sig, hist = out[rank]
        assert np.array_equal(hist, single_hist)              # partition-invariant sum
        assert np.array_equal(np.isnan(sig), np.isnan(single_sig))
        assert np.allclose(sig[~np.isnan(sig)], single_sig[~np.isnan(single_sig)])
