"""torchrun probe: ms per step of c25519_x25519_shared_sharded (mode from $C25519_SHARD_MODE / $C25519_SHARD_TAIL_DEN) vs the
plain single-GPU batch call.  Usage: torchrun --nproc-per-node N tools/shard_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from curve25519_b200 import api
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def bootstrap(uid):
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0: t.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(t, src=0); return bytes(t.cpu().numpy().tobytes())
comm = api.ShardedComm(world, rank, local, bootstrap)
n = 1 << 20
rng = np.random.Generator(np.random.PCG64(rank))
sets = [(torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda(), torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()) for _ in range(4)]
out_all = torch.empty((world * n, 32), dtype=torch.uint8, device="cuda"); out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
def timeit(fn, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
k = [0]
def sharded():
    pk, sk = sets[k[0] % 4]; k[0] += 1
    api.x25519_shared_sharded(comm, out_all, pk, sk)
def plain():
    pk, sk = sets[k[0] % 4]; k[0] += 1
    api.x25519_shared(pk, sk, out=out, sk_inplace=True)
def plain_then_gather():
    pk, sk = sets[k[0] % 4]; k[0] += 1
    api.x25519_shared(pk, sk, out=out_all[rank * n:(rank + 1) * n], sk_inplace=True)
    api.allgather_records(comm, out_all, n)
a, b, c = timeit(plain), timeit(sharded), timeit(plain_then_gather)
k[0] = 0; sharded(); torch.cuda.synchronize(); ref = out_all.clone(); out_all.zero_()
reg = comm.register(out_all)
k[0] = 0; sharded(); torch.cuda.synchronize(); dist.barrier()
same = bool(torch.equal(out_all, ref))
d, e = timeit(sharded), timeit(plain_then_gather)
comm.set_deferred(out_all, True)
def timeit_deferred(steps=10):
    for _ in range(3): sharded()
    comm.sync(out_all); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): sharded()
    comm.sync(out_all)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
f = timeit_deferred()
k[0] = 0; out_all.zero_(); torch.cuda.synchronize(); dist.barrier(); sharded(); comm.sync(out_all); torch.cuda.synchronize(); dist.barrier()
same_def = bool(torch.equal(out_all, ref))
if rank == 0:
    print("world %d deferred exchange (copy engines under the next step's ladder; last one joined): %.3f ms per step, matches NCCL: %s" % (world, f, same_def), flush=True)
if rank == 0:
    print("world %d mode %s den %s: plain %.3f ms | NCCL: sharded entry %.3f, batch + allgather %.3f | registered=%s (matches NCCL: %s): sharded entry %.3f, batch + allgather %.3f" %
          (world, os.environ.get("C25519_SHARD_MODE", "1"), os.environ.get("C25519_SHARD_TAIL_DEN", "8"), a, b, c, reg, same, d, e), flush=True)
comm.close(); dist.destroy_process_group()
