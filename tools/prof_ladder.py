"""Drive a few launches of one operation for ncu (tools/gpu_*.sh).  Usage: prof_ladder.py [n] [shared|public|ed|verify]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from curve25519_b200 import api
api.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
what = sys.argv[2] if len(sys.argv) > 2 else "shared"
rng = np.random.Generator(np.random.PCG64(1))
sk = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
pk = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
msgs = torch.from_numpy(rng.integers(0, 256, (n, 64), dtype=np.uint8)).cuda()
if what == "verify":
    pub, priv = api.ed25519_keypair(sk); sig = api.ed25519_sign(priv, msgs); sig[::16, 9] ^= 0x40
for _ in range(3):
    if what == "shared": api.x25519_shared(pk, sk)
    elif what == "public": api.x25519_public(sk)
    elif what == "verify": ok = api.ed25519_verify(sig, pub, msgs)
    elif what == "ed":
        pub, priv = api.ed25519_keypair(sk); sig = api.ed25519_sign(priv, msgs); ok = api.ed25519_verify(sig, pub, msgs)
torch.cuda.synchronize()
