"""Host-side logic of the multi-GPU path on CPU: contiguous row shards, no hot-path collective,
off-path action gather.  world_size-2 over gloo; the per-rank compute is replaced by the oracle (this tests the
sharding plumbing, not the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest

from go2_onnx_controller_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("B,world", [(0, 1), (1, 2), (7, 2), (4096, 8), (1048576, 8), (262144, 8), (13, 4), (3, 8)])
def test_shard_rows_partition(B, world):
    blocks = [shard.shard_rows(B, world, r) for r in range(world)]
    assert blocks[0][0] == 0
    for (r0, n0), (r1, _) in zip(blocks, blocks[1:]):
        assert r0 + n0 == r1
    assert blocks[-1][0] + blocks[-1][1] == B
    sizes = [n for _, n in blocks]
    assert max(sizes) - min(sizes) <= 1


def test_shard_rows_rejects_bad_arguments():
    with pytest.raises(ValueError):
        shard.shard_rows(10, 0, 0)
    with pytest.raises(ValueError):
        shard.shard_rows(10, 2, 2)
    with pytest.raises(ValueError):
        shard.shard_rows(-1, 2, 0)


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from go2_onnx_controller_b200 import DEFAULT_MODEL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pol = oracle.load_policy(DEFAULT_MODEL)
        X = oracle.make_obs_d1(B, 98, seed=0)

        def fn(row0, rows):
            return torch.from_numpy(oracle.forward(pol, X[row0:row0 + rows], np.float32).astype(np.float32))

        local = shard.run_sharded(B, rank, world, fn)
        dist.barrier()
        full = shard.gather_actions(local, B)
        ref = oracle.forward(pol, X, np.float32).astype(np.float32)
        q.put((rank, bool(np.array_equal(full.numpy(), ref)), tuple(local.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_shard_and_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    B = 1001     # odd: ranks get 501 / 500 rows
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res[0] == (0, True, (501, 12)) and res[1] == (1, True, (500, 12))
