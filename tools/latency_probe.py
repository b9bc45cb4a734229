"""n = 1 and small-batch latency of the host-pointer ABI (legacy wrappers included).  Usage: python tools/latency_probe.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from curve25519_b200 import _native, api
L = _native.lib()
api.init(0)
sk = (C.c_uint8 * 32)(*range(32)); pk = (C.c_uint8 * 32)(9); out = (C.c_uint8 * 32)()
def bench(fn, reps=300):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e6
print("curve25519_dh_CreateSharedKey (n = 1, ladder)      %8.1f us" % bench(lambda: L.curve25519_dh_CreateSharedKey(out, pk, sk)))
print("curve25519_dh_CalculatePublicKey (n = 1, ladder)   %8.1f us" % bench(lambda: L.curve25519_dh_CalculatePublicKey(out, sk)))
print("curve25519_dh_CalculatePublicKey_fast (n = 1, comb)%8.1f us" % bench(lambda: L.curve25519_dh_CalculatePublicKey_fast(out, sk)))
rng = np.random.Generator(np.random.PCG64(3))
for n in (1, 8, 32, 128, 255, 256, 1024, 4096):
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8); p = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    us = bench(lambda: api.x25519_shared(p, s), 100)
    print("c25519_x25519_shared_host n = %5d                 %8.1f us  (%.2f us/op)" % (n, us, us / n))
