#!/usr/bin/env python3
"""Generate tests/golden/*.json from the COMPILED REFERENCE (oracle/_ref/libref25519.so, built from
/root/reference by oracle/Makefile).  Run in the authoring container only; the JSON files are committed
so that boxes without the reference can still pin the restatement and the CUDA engine.

    python tests/golden/gen_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.pyoracle import Oracle  # noqa: E402

R = Oracle("reference")
rng = np.random.Generator(np.random.PCG64(0xC25519))
hexrows = lambda a: [r.tobytes().hex() for r in a]

# ---- X25519: random records + hand-picked degenerate u-coordinates and scalars
n = 96
sk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
special_u = ["00" * 32, "01" + "00" * 31, "ec" + "ff" * 30 + "7f", "ed" + "ff" * 30 + "7f", "ee" + "ff" * 30 + "7f",
             "ff" * 32, "09" + "00" * 30 + "80", "da" + "ff" * 31, "db" + "ff" * 31,       # 2^256-38 (== 0), 2^256-37
             "e0eb7a7c3b41b8ae1656e3faf19fc46ada098deb9c32b1fd866205165f49b800",
             "5f9c95bca3508c24b1d0b1559c83ef5b04445cc4581c8e86d8224eddd09f1157", "09" + "00" * 31]
for i, u in enumerate(special_u):
    pk[i] = np.frombuffer(bytes.fromhex(u), np.uint8)
sk[20] = 0; sk[21] = 0xFF; sk[22] = np.arange(32, dtype=np.uint8)
shared, skc = R.x25519_shared(pk, sk)
public, _ = R.x25519_public(sk, fast=True)
public2, _ = R.x25519_public(sk, fast=False)
assert (public == public2).all()
json.dump({"sk": hexrows(sk), "pk": hexrows(pk), "shared": hexrows(shared), "sk_clamped": hexrows(skc), "public": hexrows(public)},
          open(os.path.join(HERE, "x25519.json"), "w"), indent=0)

# ---- Ed25519: ragged messages (0..300 bytes incl. SHA-512 padding boundaries), tampered signatures, S+L, garbage
n = 64
seed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
lens = list(rng.integers(0, 300, n))
for i, l in enumerate([0, 1, 15, 16, 17, 47, 48, 49, 63, 64, 65, 79, 80, 81, 111, 112, 113, 127, 128, 129, 143, 144, 145, 175, 176, 177, 239, 240, 241]):
    lens[i] = l
msgs = [rng.integers(0, 256, int(l), dtype=np.uint8).tobytes() for l in lens]
off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum([len(m) for m in msgs])
flat = np.frombuffer(b"".join(msgs), np.uint8)
pub, priv = R.ed25519_keypair(seed)
sig = R.ed25519_sign(priv, flat, off)
assert R.ed25519_verify(sig, pub, flat, off).all()
L = 2**252 + 27742317777372353535851937790883648493
tam = sig.copy()
for i in range(n):
    if i % 4 == 0:
        tam[i, (7 * i) % 64] ^= 1 << (i % 8)
    elif i % 4 == 1:                                  # S + L where it still fits in 256 bits -> still valid
        s = int.from_bytes(tam[i, 32:].tobytes(), "little") + L
        if s < 2**256:
            tam[i, 32:] = np.frombuffer(s.to_bytes(32, "little"), np.uint8)
ok = R.ed25519_verify(tam, pub, flat, off)
json.dump({"seed": hexrows(seed), "pub": hexrows(pub), "msg": [m.hex() for m in msgs], "sig": hexrows(sig),
           "sig_tampered": hexrows(tam), "ok_tampered": ok.tolist()},
          open(os.path.join(HERE, "ed25519.json"), "w"), indent=0)
print("golden fixtures written:", n, "ed25519 items,", 96, "x25519 items; valid after tamper:", int(ok.sum()))
