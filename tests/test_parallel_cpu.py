"""CPU: the multi-GPU host logic (row sharding + the single label all-gather) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from traffic_classifier_sdn_b200.parallel import predict_sharded, shard_bounds


def test_shard_bounds_cover_and_partition():
    for n in (0, 1, 7, 8, 9, 1000, 1001):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert all(0 <= b - a <= -(-n // world) for a, b in blocks)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


class _ParityModel:
    """stands in for an estimator: label = (first feature + row count seen by this rank) mod 5 would depend on the
    shard, so use a pure function of the row instead"""
    def predict_indices(self, X):
        X = np.asarray(X)
        return (X[:, 0].astype(np.int64) % 5).astype(np.int32)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = np.arange(n * 3, dtype=np.float64).reshape(n, 3)
        full = predict_sharded(_ParityModel(), X, gather=True)
        mine = predict_sharded(_ParityModel(), X, gather=False)
        q.put((rank, full.tolist(), mine.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11, 1])
def test_predict_sharded_gloo_world2(n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs: p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    exp = ((np.arange(n) * 3) % 5).astype(np.int32).tolist()
    for rank, full, mine in got:
        assert full == exp
        a, b = shard_bounds(n, rank, 2)
        assert mine == exp[a:b]
