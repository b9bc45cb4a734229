"""GPU parity for Ed25519 keygen / sign / verify (single- and two-phase) through the C ABI, against the CPU
checkers, RFC 8032 and the committed golden fixtures.  Bit-exact: public keys, private-key records,
signatures and verdicts must be identical, including the reference's permissive behaviours (S + L accepted,
no point validation, deterministic result on undecodable keys)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import vectors as V
from tests.conftest import hx

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NCPU = os.cpu_count() or 1


def _rows(hexes):
    return np.stack([hx(h) for h in hexes])


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ragged(msgs):
    off = np.zeros(len(msgs) + 1, np.uint64); off[1:] = np.cumsum([len(m) for m in msgs])
    return np.frombuffer(b"".join(msgs), np.uint8).copy(), off


def test_device_primitives_sha512_and_mod_l(engine, rng):
    """op 9: SHA-512 of a 64-byte string; op 8: 512-bit value mod L -- against hashlib / Python integers."""
    import hashlib
    import torch
    n = 2048
    a = rng.integers(0, 256, (n, 32), dtype=np.uint8); b = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    a[0] = 0; b[0] = 0; a[1] = 255; b[1] = 255
    # values straddling multiples of L
    for j, k in enumerate([1, 2, 15, 16, 17, 2**255 // V.L_ORDER, (2**512 - 1) // V.L_ORDER]):
        for d in (-1, 0, 1):
            v = k * V.L_ORDER + d
            raw = v.to_bytes(64, "little"); a[10 + 3 * j + d + 1] = np.frombuffer(raw[:32], np.uint8); b[10 + 3 * j + d + 1] = np.frombuffer(raw[32:], np.uint8)
    dg = engine.test_primitive(9, _dev(a), _dev(b), out_rec=64).cpu().numpy()
    red = engine.test_primitive(8, _dev(a), _dev(b)).cpu().numpy()
    for i in range(n):
        raw = a[i].tobytes() + b[i].tobytes()
        assert dg[i].tobytes() == hashlib.sha512(raw).digest(), i
        assert int.from_bytes(red[i].tobytes(), "little") == int.from_bytes(raw, "little") % V.L_ORDER, i


def test_rfc8032_vectors(engine):
    for seed, pk, msg, sig in V.ED25519_KAT:
        pub, priv = engine.ed25519_keypair(_dev(hx(seed)[None, :]))
        assert pub.cpu().numpy()[0].tobytes().hex() == pk
        assert priv.cpu().numpy()[0].tobytes().hex() == seed + pk
        flat, off = _ragged([bytes.fromhex(msg)])
        import torch
        d_flat = _dev(flat) if flat.size else torch.zeros(1, dtype=torch.uint8, device="cuda")
        d_off = torch.from_numpy(off.astype(np.int64)).cuda()
        s = engine.ed25519_sign(priv, d_flat, d_off)
        assert s.cpu().numpy()[0].tobytes().hex() == sig
        assert engine.ed25519_verify(s, pub, d_flat, d_off).cpu().numpy()[0] == 1
        # host-pointer flavour
        s_h = engine.ed25519_sign(priv.cpu().numpy(), flat, off)
        assert s_h[0].tobytes().hex() == sig
        assert engine.ed25519_verify(s_h, pub.cpu().numpy(), flat, off)[0] == 1


def test_golden_fixture(engine):
    import torch
    g = json.load(open(os.path.join(GOLD, "ed25519.json")))
    seed = _rows(g["seed"])
    pub, priv = engine.ed25519_keypair(_dev(seed))
    assert [r.tobytes().hex() for r in pub.cpu().numpy()] == g["pub"]
    flat, off = _ragged([bytes.fromhex(m) for m in g["msg"]])
    d_flat, d_off = _dev(flat), torch.from_numpy(off.astype(np.int64)).cuda()
    sig = engine.ed25519_sign(priv, d_flat, d_off)
    assert [r.tobytes().hex() for r in sig.cpu().numpy()] == g["sig"]
    ok = engine.ed25519_verify(_dev(_rows(g["sig_tampered"])), pub, d_flat, d_off)
    assert ok.cpu().numpy().tolist() == g["ok_tampered"]


@pytest.mark.parametrize("n,mlen", [(1, 0), (33, 1), (129, 64), (1000, 64), (257, 111), (257, 112), (64, 239), (5000, 64)])
def test_fixed_length_batches_vs_oracle(engine, oracle, rng, n, mlen):
    """keygen + sign + verify on seeded random batches; 1/4 of the signatures are corrupted in R, S or the
    message so that both verdicts occur."""
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, mlen), dtype=np.uint8)
    exp_pub, exp_priv = oracle.ed25519_keypair(seed, threads=NCPU)
    pub, priv = engine.ed25519_keypair(_dev(seed))
    assert (pub.cpu().numpy() == exp_pub).all() and (priv.cpu().numpy() == exp_priv).all()
    exp_sig = oracle.ed25519_sign(exp_priv, msgs, threads=NCPU)
    sig = engine.ed25519_sign(priv, _dev(msgs))
    assert (sig.cpu().numpy() == exp_sig).all()
    bad = exp_sig.copy(); bmsgs = msgs.copy()
    for i in range(0, n, 4):
        which = (i // 4) % 3
        if which == 0: bad[i, (5 * i) % 32] ^= 1 << (i % 8)
        elif which == 1: bad[i, 32 + (3 * i) % 32] ^= 1 << (i % 8)
        elif mlen: bmsgs[i, i % mlen] ^= 1
    exp_ok = oracle.ed25519_verify(bad, exp_pub, bmsgs, threads=NCPU)
    ok = engine.ed25519_verify(_dev(bad), pub, _dev(bmsgs))
    assert (ok.cpu().numpy() == exp_ok).all()
    assert 0 < exp_ok.sum() < n or n < 4
    # host-pointer flavour of all three
    pub_h, priv_h = engine.ed25519_keypair(seed)
    assert (pub_h == exp_pub).all() and (priv_h == exp_priv).all()
    assert (engine.ed25519_sign(exp_priv, msgs) == exp_sig).all()
    assert (engine.ed25519_verify(bad, exp_pub, bmsgs) == exp_ok).all()


def test_ragged_messages_vs_oracle(engine, oracle, rng):
    import torch
    n = 777
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    lens = rng.integers(0, 400, n)
    lens[:12] = [0, 1, 47, 48, 49, 111, 112, 113, 175, 176, 177, 399]
    flat = rng.integers(0, 256, int(lens.sum()), dtype=np.uint8)
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    pub, priv = oracle.ed25519_keypair(seed, threads=NCPU)
    exp_sig = oracle.ed25519_sign(priv, flat, off, threads=NCPU)
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    sig = engine.ed25519_sign(_dev(priv), _dev(flat), d_off)
    assert (sig.cpu().numpy() == exp_sig).all()
    bad = exp_sig.copy(); bad[::5, 7] ^= 0x10
    exp_ok = oracle.ed25519_verify(bad, pub, flat, off, threads=NCPU)
    assert (engine.ed25519_verify(_dev(bad), _dev(pub), _dev(flat), d_off).cpu().numpy() == exp_ok).all()
    assert (engine.ed25519_sign(priv, flat, off) == exp_sig).all()
    assert (engine.ed25519_verify(bad, pub, flat, off) == exp_ok).all()


def test_permissive_verification_quirks(engine, oracle, rng):
    """S + L verifies; undecodable / random public keys never crash and give the checker's verdict."""
    n = 256
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8); msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    pub, priv = oracle.ed25519_keypair(seed, threads=NCPU)
    sig = oracle.ed25519_sign(priv, msgs, threads=NCPU)
    forged = sig.copy(); cnt = 0
    for i in range(n):
        s = int.from_bytes(sig[i, 32:].tobytes(), "little") + V.L_ORDER * (1 + i % 15)
        if s < 2**256:
            forged[i, 32:] = np.frombuffer(s.to_bytes(32, "little"), np.uint8); cnt += 1
    exp = oracle.ed25519_verify(forged, pub, msgs, threads=NCPU)
    assert exp.sum() == n and cnt > n // 2
    assert (engine.ed25519_verify(_dev(forged), _dev(pub), _dev(msgs)).cpu().numpy() == exp).all()
    garbage = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    garbage[0] = 0; garbage[1] = 255; garbage[2] = hx("01" + "00" * 31); garbage[3] = hx("ec" + "ff" * 30 + "7f")
    exp = oracle.ed25519_verify(sig, garbage, msgs, threads=NCPU)
    assert (engine.ed25519_verify(_dev(sig), _dev(garbage), _dev(msgs)).cpu().numpy() == exp).all()
    # non-canonical encodings of y (y + p) in the public key and high bit games in R
    nc_pub = pub.copy()
    exp = oracle.ed25519_verify(sig, nc_pub, msgs, threads=NCPU)
    assert (engine.ed25519_verify(_dev(sig), _dev(nc_pub), _dev(msgs)).cpu().numpy() == exp).all()


def test_two_phase_verification(engine, oracle, rng):
    """ed25519_Verify_Init once per key, ed25519_Verify_Check for many signatures (ed25519_verify.c:282-286)."""
    import torch
    nk, per = 7, 40
    seed = rng.integers(0, 256, (nk, 32), dtype=np.uint8)
    pub, priv = oracle.ed25519_keypair(seed)
    n = nk * per
    kidx = rng.integers(0, nk, n).astype(np.int32)
    msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    sig = oracle.ed25519_sign(priv[kidx], msgs, threads=NCPU)
    sig[::6, 3] ^= 2
    exp = oracle.ed25519_verify(sig, pub[kidx], msgs, threads=NCPU)
    ctx = engine.ed25519_verify_init(_dev(pub))
    assert tuple(ctx.shape) == (nk, 2080)
    ok = engine.ed25519_verify_check(ctx, _dev(sig), _dev(msgs), key_index=torch.from_numpy(kidx).cuda())
    assert (ok.cpu().numpy() == exp).all() and 0 < exp.sum() < n


def test_legacy_wrappers_signature_test(engine):
    """The reference's signature_test (test/curve25519_test.c:323-410): RFC 8032 TEST 2 through the
    re-exported n = 1 symbols, with and without a blinding context, single- and two-phase verify."""
    from curve25519_b200 import _native
    L = _native.lib()
    seed, pk, msg, sig = V.ED25519_KAT[1]
    sk = (C.c_uint8 * 32).from_buffer_copy(bytes.fromhex(seed)); m = (C.c_uint8 * 1).from_buffer_copy(bytes.fromhex(msg))
    pub = (C.c_uint8 * 32)(); priv = (C.c_uint8 * 64)(); s = (C.c_uint8 * 64)()
    L.ed25519_CreateKeyPair(pub, priv, None, sk)
    assert bytes(pub).hex() == pk and bytes(priv).hex() == seed + pk
    L.ed25519_SignMessage(s, priv, None, m, 1)
    assert bytes(s).hex() == sig
    assert L.ed25519_VerifySignature(s, pub, m, 1) == 1
    blind = L.ed25519_Blinding_Init(None, bytes(range(64)), 64)
    s2 = (C.c_uint8 * 64)(); pub2 = (C.c_uint8 * 32)(); priv2 = (C.c_uint8 * 64)()
    L.ed25519_CreateKeyPair(pub2, priv2, blind, sk)
    L.ed25519_SignMessage(s2, priv2, blind, m, 1)
    L.ed25519_Blinding_Finish(blind)
    assert bytes(s2).hex() == sig and bytes(pub2).hex() == pk
    ctx = L.ed25519_Verify_Init(None, pub)
    assert L.ed25519_Verify_Check(ctx, s, m, 1) == 1
    s[0] ^= 1
    assert L.ed25519_Verify_Check(ctx, s, m, 1) == 0
    assert L.ed25519_VerifySignature(s, pub, m, 1) == 0
    L.ed25519_Verify_Finish(ctx)
    store = (C.c_uint8 * 2080)()            # caller-supplied storage of the reference's sizeof(EDP_SIGV_CTX)
    s[0] ^= 1
    assert L.ed25519_Verify_Init(store, pub) == C.addressof(store)
    assert L.ed25519_Verify_Check(store, s, m, 1) == 1


def test_full_size_1m_round_trip(engine, oracle, rng):
    """BASELINE config 4 at full size: 2^20 keygen -> sign -> verify round trip; a deterministic 1/16 of the
    signatures is corrupted; the verdict vector must be exactly the corruption mask, and 2048 sampled
    records must equal the oracle byte for byte."""
    import torch
    n = 1 << 20
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8); msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    d_seed, d_msgs = _dev(seed), _dev(msgs)
    pub, priv = engine.ed25519_keypair(d_seed)
    sig = engine.ed25519_sign(priv, d_msgs)
    mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[::16] = True
    bad = sig.clone(); bad[::16, 9] ^= 0x40
    ok = engine.ed25519_verify(bad, pub, d_msgs)
    assert torch.equal(ok == 0, mask)
    # whole batch against the oracle when the host has the cores for it, else a sample
    idx = np.arange(n) if NCPU >= 32 else rng.choice(n, 2048, replace=False)
    exp_pub, exp_priv = oracle.ed25519_keypair(seed[idx], threads=NCPU)
    assert (pub.cpu().numpy()[idx] == exp_pub).all() and (priv.cpu().numpy()[idx] == exp_priv).all()
    exp_sig = oracle.ed25519_sign(exp_priv, msgs[idx], threads=NCPU)
    assert (sig.cpu().numpy()[idx] == exp_sig).all()
    exp_ok = oracle.ed25519_verify(bad.cpu().numpy()[idx], exp_pub, msgs[idx], threads=NCPU)
    assert (ok.cpu().numpy()[idx] == exp_ok).all()


def test_device_sc_muladd(engine, rng):
    """op 11: (a*b + a) mod L through sc_muladd, incl. operands >= L and all-ones."""
    n = 2048
    a = rng.integers(0, 256, (n, 32), dtype=np.uint8); b = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    a[0] = 255; b[0] = 255; a[1] = 0; b[2] = 0
    for j, v in enumerate([V.L_ORDER, V.L_ORDER - 1, V.L_ORDER + 1, 15 * V.L_ORDER, 2**252, 2**256 - 1]):
        a[10 + j] = np.frombuffer(v.to_bytes(32, "little"), np.uint8); b[20 + j] = a[10 + j]
    got = engine.test_primitive(11, _dev(a), _dev(b)).cpu().numpy()
    for i in range(n):
        x = int.from_bytes(a[i].tobytes(), "little"); y = int.from_bytes(b[i].tobytes(), "little")
        assert int.from_bytes(got[i].tobytes(), "little") == (x * y + x) % V.L_ORDER, i


def test_long_messages(engine, oracle, rng):
    """Multi-block hashing on the device: messages of 1 KB .. 64 KB (up to 513 SHA-512 blocks), unaligned starts."""
    import torch
    lens = [1000, 4096, 4097, 65535, 65536, 12345, 8191, 3]
    n = len(lens)
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    pub, priv = oracle.ed25519_keypair(seed)
    flat = rng.integers(0, 256, int(sum(lens)), dtype=np.uint8)
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    exp = oracle.ed25519_sign(priv, flat, off)
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    sig = engine.ed25519_sign(_dev(priv), _dev(flat), d_off)
    assert (sig.cpu().numpy() == exp).all()
    bad = exp.copy(); bad[1, 0] ^= 1
    ok = engine.ed25519_verify(_dev(bad), _dev(pub), _dev(flat), d_off).cpu().numpy()
    assert ok.tolist() == [1, 0, 1, 1, 1, 1, 1, 1]
    assert (engine.ed25519_sign(priv, flat, off) == exp).all()


def test_ragged_host_path_is_sliced(engine, oracle, rng):
    """Ragged messages through the HOST-pointer ABI with more operations than one pipeline slice holds (2^17): the slices are
    staged with unmodified offsets and a message base pointer shifted back by the slice's first offset; results must equal the
    device-pointer path on the same inputs, and a sample must equal the oracle."""
    import torch
    n = (1 << 17) * 2 + 12345
    lens = rng.integers(0, 41, n)
    lens[: 16] = [0, 1, 15, 16, 17, 40, 0, 0, 39, 2, 3, 4, 5, 6, 7, 8]
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    flat = rng.integers(0, 256, int(off[-1]), dtype=np.uint8)
    seed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    pub, priv = engine.ed25519_keypair(seed)                       # host path (fixed-size records)
    sig_h = engine.ed25519_sign(priv, flat, off)                   # host path, ragged: three slices
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    sig_d = engine.ed25519_sign(_dev(priv), _dev(flat), d_off).cpu().numpy()
    assert (sig_h == sig_d).all()
    bad = sig_h.copy(); bad[::7, 33] ^= 2
    ok_h = engine.ed25519_verify(bad, pub, flat, off)
    ok_d = engine.ed25519_verify(_dev(bad), _dev(pub), _dev(flat), d_off).cpu().numpy()
    assert (ok_h == ok_d).all() and (ok_h[::7] == 0).all() and ok_h.sum() == n - len(range(0, n, 7))
    idx = np.concatenate([np.arange(64), np.arange((1 << 17) - 32, (1 << 17) + 32), np.arange(n - 64, n)])   # around the slice boundaries
    s_off = np.zeros(idx.size + 1, np.uint64); s_off[1:] = np.cumsum(lens[idx])
    s_flat = np.concatenate([flat[int(off[i]):int(off[i + 1])] for i in idx]) if s_off[-1] else np.zeros(0, np.uint8)
    e_pub, e_priv = oracle.ed25519_keypair(seed[idx])
    assert (e_pub == pub[idx]).all()
    assert (oracle.ed25519_sign(e_priv, s_flat, s_off) == sig_h[idx]).all()


def test_empty_and_misaligned_batches(engine):
    """n = 0 is a successful no-op for every batched entry point (also with null pointers); misaligned device record arrays and
    host pointers handed to *_batch are refused before anything is launched."""
    import ctypes as C
    import torch
    from curve25519_b200 import _native
    L = _native.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    before = L.c25519_launch_count()
    assert L.c25519_ed25519_keypair_batch(None, None, None, 0, s) == 0
    assert L.c25519_ed25519_sign_batch(None, None, None, None, 0, 0, s) == 0
    assert L.c25519_ed25519_verify_batch(None, None, None, None, None, 0, 0, s) == 0
    assert L.c25519_x25519_public_batch(None, None, 0, 0, s) == 0
    assert L.c25519_modl_batch(0, None, None, None, 0, s) == 0
    assert L.c25519_ed25519_verify_host(None, None, None, None, None, 0, 0) == 0
    assert L.c25519_x25519_shared_host(None, None, None, 0) == 0
    buf = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    base = buf.data_ptr() + (-buf.data_ptr()) % 32
    assert L.c25519_ed25519_keypair_batch(C.c_void_p(base + 4), C.c_void_p(base + 64), C.c_void_p(base + 128), 1, s) == -3
    assert L.c25519_modl_batch(9, C.c_void_p(base), C.c_void_p(base + 32), C.c_void_p(base + 64), 1, s) == -3
    host = np.zeros((4, 32), np.uint8)
    hp = C.c_void_p(host.ctypes.data + (-host.ctypes.data) % 32)
    rc = L.c25519_x25519_public_batch(hp, hp, 1, 0, s)
    assert rc != 0 and L.c25519_last_error()
    assert L.c25519_launch_count() == before
