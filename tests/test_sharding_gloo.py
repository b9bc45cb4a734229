"""CPU, world_size 2 (and 3), gloo: the N > 1 host logic -- contiguous sharding, ragged tail, padding, the
single all-gather of result records -- with the CPU checker standing in for the kernels (no GPU here).
The gathered buffer on EVERY rank must equal the unsharded oracle output."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from curve25519_b200 import sharding
    from oracle.pyoracle import Oracle
    o = Oracle("port")
    rng = np.random.Generator(np.random.PCG64(99))
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    full, _ = o.x25519_shared(pk, sk) if n else (np.zeros((0, 32), np.uint8), None)

    def fn(pk_s, sk_s):          # stands in for api.x25519_shared on the rank's GPU
        out, _ = o.x25519_shared(pk_s.numpy(), sk_s.numpy())
        return torch.from_numpy(out)

    got = sharding.sharded_apply(fn, n, [torch.from_numpy(pk), torch.from_numpy(sk)], 32)
    ok = bool((got.numpy() == full).all()) and tuple(got.shape) == (n, 32)
    # verdict-style 1-D int32 records through the same collective
    lo, hi = sharding.shard_bounds(n, world, rank)
    v = torch.arange(lo, hi, dtype=torch.int32)
    g = sharding.all_gather_records(v, n)
    ok = ok and bool((g == torch.arange(n, dtype=torch.int32)).all())
    # bounds tile [0, n) exactly
    cover = []
    for r in range(world):
        a, b = sharding.shard_bounds(n, world, r); cover += list(range(a, b))
    ok = ok and cover == list(range(n))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 64), (2, 65), (2, 1), (3, 10), (2, 0)])
def test_sharded_allgather_matches_unsharded(world, n):
    port = _free_port()
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
