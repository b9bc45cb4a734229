#!/usr/bin/env python3
"""Generate tests/golden/fullsize_digests.json from the COMPILED REFERENCE (oracle/_ref/libref25519.so).

Runs the reference's n = 1 API over the full 2^20-record inputs of BASELINE configs 2, 3 and 4
(tests/fullsize.py) on all host cores (about two minutes on 8 cores) and stores SHA-256 digests of every
output array (whole array + per 4096-record block).  Authoring container only; the JSON is committed.

    python tests/golden/gen_digests.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import fullsize as F  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

R = Oracle("reference")
T = os.cpu_count() or 1
out = {"n": F.N_FULL, "block": F.BLOCK, "generator": "oracle/_ref/libref25519.so (reference portable-C), tests/golden/gen_digests.py"}
t0 = time.time()

sk, pk = F.config2_inputs()
shared, skc = R.x25519_shared(pk, sk, threads=T)
out["config2_shared"] = F.digests(shared)
out["config2_sk_clamped"] = F.digests(skc)
print("config 2 done %.0f s" % (time.time() - t0), flush=True)
public, _ = R.x25519_public(sk, fast=True, threads=T)
out["config3_public"] = F.digests(public)
print("config 3 done %.0f s" % (time.time() - t0), flush=True)

seed, msgs = F.config4_inputs()
pub, priv = R.ed25519_keypair(seed, threads=T)
sig = R.ed25519_sign(priv, msgs, threads=T)
out["config4_pub"] = F.digests(pub)
out["config4_priv"] = F.digests(priv)
out["config4_sig"] = F.digests(sig)
tsig, tmsgs = F.config4_tamper(sig, msgs)
ok = R.ed25519_verify(tsig, pub, tmsgs, threads=T)
out["config4_ok"] = F.digests(ok)
out["config4_ok_count"] = int(ok.sum())
print("config 4 done %.0f s; valid %d of %d" % (time.time() - t0, int(ok.sum()), F.N_FULL), flush=True)
json.dump(out, open(os.path.join(HERE, "fullsize_digests.json"), "w"), indent=0)
