"""CPU-only: the full-size fixture (tests/fullsize.py + tests/golden/fullsize_digests.json) is self-consistent, and the
oracle reproduces the committed digests on a prefix of every configuration (the whole 2^20-record arrays are what the GPU
tests check; here the first 4096-record block of each, which is seconds of CPU work)."""
import json
import os

import numpy as np

from tests import fullsize as F

HERE = os.path.dirname(os.path.abspath(__file__))


def test_inputs_are_deterministic_and_tamper_is_sparse():
    sk, pk = F.config2_inputs(8192); sk2, pk2 = F.config2_inputs(8192)
    assert (sk == sk2).all() and (pk == pk2).all() and not (sk == pk).all()
    # a prefix of the full-size inputs is the small-size inputs (PCG64 streams row by row per array)
    big_sk, _ = F.config2_inputs(16384)
    assert (big_sk[:8192] == sk).all()
    seed, msgs = F.config4_inputs(4096)
    sig = np.zeros((4096, 64), np.uint8)
    tsig, tmsgs = F.config4_tamper(sig, msgs)
    changed = np.flatnonzero((tsig != sig).any(axis=1) | (tmsgs != msgs).any(axis=1))
    flipped = changed[changed % 16 == 0]
    assert len(flipped) == 4096 // 16                       # exactly one bit in every 16th item
    for i in flipped:
        bits = int(np.unpackbits(tsig[i] ^ sig[i]).sum() + np.unpackbits(tmsgs[i] ^ msgs[i]).sum())
        assert bits == 1, i
    assert set(changed) - set(flipped) <= set(range(8, 64, 16))   # the S + L items (only where S + L < 2^256)


def test_digest_file_shape():
    d = json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))
    assert d["n"] == F.N_FULL and d["block"] == F.BLOCK
    for k in ("config2_shared", "config2_sk_clamped", "config3_public", "config4_pub", "config4_priv", "config4_sig", "config4_ok"):
        assert len(d[k]["blocks"]) == F.N_FULL // F.BLOCK and len(d[k]["all"]) == 64, k
    assert d["config4_ok_count"] == F.N_FULL - F.N_FULL // 16


def test_oracle_reproduces_the_first_block_of_every_digest(oracle):
    d = json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))
    T = os.cpu_count() or 1
    n = F.BLOCK
    sk, pk = F.config2_inputs()
    out, skc = oracle.x25519_shared(pk[:n], sk[:n], threads=T)
    assert F.digests(out)["blocks"][0] == d["config2_shared"]["blocks"][0]
    assert F.digests(skc)["blocks"][0] == d["config2_sk_clamped"]["blocks"][0]
    pub, _ = oracle.x25519_public(sk[:n], fast=True, threads=T)
    assert F.digests(pub)["blocks"][0] == d["config3_public"]["blocks"][0]
    seed, msgs = F.config4_inputs()
    epub, epriv = oracle.ed25519_keypair(seed[:n], threads=T)
    sig = oracle.ed25519_sign(epriv, msgs[:n], threads=T)
    assert F.digests(epub)["blocks"][0] == d["config4_pub"]["blocks"][0]
    assert F.digests(sig)["blocks"][0] == d["config4_sig"]["blocks"][0]
    tsig, tmsgs = F.config4_tamper(sig, msgs[:n])
    ok = oracle.ed25519_verify(tsig, epub, tmsgs, threads=T)
    assert F.digests(ok)["blocks"][0] == d["config4_ok"]["blocks"][0]
