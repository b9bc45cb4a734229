"""Deterministic full-size inputs of BASELINE configs 2, 3, 4 and block digests of outputs.

The GPU box has neither /root/reference nor the cores to run the oracle over 2^20 records in a test, so 100 % byte
equality at full size is checked through SHA-256 digests committed in tests/golden/fullsize_digests.json:
tests/golden/gen_digests.py runs the COMPILED REFERENCE (oracle/_ref) over exactly these inputs in the authoring
container and stores, per output array, one digest per block of 4096 records plus the digest of the whole array; the
`-m gpu` tests recompute the same digests from the CUDA engine's outputs.  A mismatch names the 4096-record block.

numpy only -- no oracle, no engine: both sides import this module for the inputs.
"""
import hashlib

import numpy as np

N_FULL = 1 << 20
BLOCK = 4096
L_ORDER = 2**252 + 27742317777372353535851937790883648493


def config2_inputs(n=N_FULL):
    """Config 2/3: uniform random 32-byte scalars then points from PCG64(0x25519); no clamping, no bit-255 masking."""
    rng = np.random.Generator(np.random.PCG64(0x25519))
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    return sk, pk


def config4_inputs(n=N_FULL):
    """Config 4: random 32-byte seeds and 64-byte messages from PCG64(0xED25519)."""
    rng = np.random.Generator(np.random.PCG64(0xED25519))
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    return seed, msgs


def config4_tamper(sig, msgs):
    """Corrupt a deterministic subset in place-free fashion (returns copies): for i % 16 == 0 flip ONE bit of
    R || S || msg (1024 bits) at a position hashed from i; for i % 16 == 8 and i % 4096 < 64 replace S by S + L when that
    still fits 256 bits (the reference accepts it, ed25519_verify.c:308 uses S raw).  -> (sig', msgs')."""
    sig = sig.copy(); msgs = msgs.copy()
    n = sig.shape[0]
    idx = np.arange(0, n, 16, dtype=np.uint64)
    h = (idx * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(20)
    pos = (h % np.uint64(1024)).astype(np.int64)
    byte, bit = pos >> 3, (pos & 7).astype(np.uint8)
    rows = idx.astype(np.int64)
    in_sig = byte < 64
    sig[rows[in_sig], byte[in_sig]] ^= (np.uint8(1) << bit[in_sig])
    msgs[rows[~in_sig], byte[~in_sig] - 64] ^= (np.uint8(1) << bit[~in_sig])
    for i in range(8, n, 16):
        if i % 4096 < 64:
            s = int.from_bytes(sig[i, 32:].tobytes(), "little") + L_ORDER
            if s < 2**256:
                sig[i, 32:] = np.frombuffer(s.to_bytes(32, "little"), np.uint8)
    return sig, msgs


def digests(a):
    """-> {"all": sha256 of the whole array, "blocks": [sha256 of each 4096-record block]}"""
    a = np.ascontiguousarray(a)
    flat = a.reshape(a.shape[0], -1).view(np.uint8)
    blocks = [hashlib.sha256(flat[i:i + BLOCK].tobytes()).hexdigest() for i in range(0, flat.shape[0], BLOCK)]
    return {"all": hashlib.sha256(flat.tobytes()).hexdigest(), "blocks": blocks}


def assert_digest(name, got_array, expected):
    d = digests(got_array)
    if d["all"] == expected["all"]:
        return
    bad = [i for i, (x, y) in enumerate(zip(d["blocks"], expected["blocks"])) if x != y]
    raise AssertionError("%s differs from the reference in %d of %d blocks of %d records (first: block %d)"
                         % (name, len(bad), len(expected["blocks"]), BLOCK, bad[0] if bad else -1))
