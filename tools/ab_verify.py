"""A/B timing of the two verify_check loop variants in ONE process (alternating), tables cached."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for rep in range(2):
        for v in ("split", "unified"):
            env = dict(os.environ)
            if v == "unified": env["C25519_VERIFY_UNIFIED_ADD"] = "1"
            else: env.pop("C25519_VERIFY_UNIFIED_ADD", None)
            print(v, subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
    sys.exit(0)
import numpy as np, torch
from curve25519_b200 import api
api.init(0)
n = 1 << 20
rng = np.random.Generator(np.random.PCG64(3))
seed = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
msgs = torch.from_numpy(rng.integers(0, 256, (n, 64), dtype=np.uint8)).cuda()
pub, priv = api.ed25519_keypair(seed); sig = api.ed25519_sign(priv, msgs)
def t(fn, k=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k
ms = t(lambda: api.ed25519_verify(sig, pub, msgs))
ctx = api.ed25519_verify_init(pub[:4096].contiguous()); kidx = (torch.arange(n, device="cuda", dtype=torch.int32) % 4096).contiguous()
sig_k = api.ed25519_sign(priv[:4096].contiguous()[kidx.long()].contiguous(), msgs)
ms2 = t(lambda: api.ed25519_verify_check(ctx, sig_k, msgs, key_index=kidx))
ok = bool(api.ed25519_verify(sig, pub, msgs).all())
print("verify %.3f ms (%.1f M/s)  check-cached %.3f ms (%.1f M/s) ok=%s" % (ms, n / ms / 1e3, ms2, n / ms2 / 1e3, ok))
