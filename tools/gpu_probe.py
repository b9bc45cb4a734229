"""First-contact GPU probe: times the hot kernels with CUDA events and prints the IMAD peak."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from curve25519_b200 import api
api.init(0)
peak = api.imad_peak()
print("imad_peak (fresh form) MAC32/s = %.4e" % peak)
rng = np.random.Generator(np.random.PCG64(1))
res = {}
for n in (1 << 16, 1 << 18, 1 << 20, 1 << 22):
    sk = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
    pk = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
    out = torch.empty_like(sk)
    for _ in range(2): api.x25519_shared(pk, sk, out=out, sk_inplace=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): api.x25519_shared(pk, sk, out=out, sk_inplace=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ops = n / (ms * 1e-3)
    print("x25519_shared n=%d  %.3f ms  %.3e ops/s  algorithmic MAC frac=%.3f" % (n, ms, ops, ops * 150783 / peak))
    res[n] = ops
json.dump({"imad_peak": peak, "x25519_shared_ops": res}, open("gpurun_out/probe.json", "w"))
