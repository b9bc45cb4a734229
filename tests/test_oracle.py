"""CPU-only: pins the oracle.  (1) the C restatement and the compiled reference against the RFC 7748 /
RFC 8032 vectors and the reference's own KATs; (2) the restatement against the compiled reference on
seeded random and degenerate inputs; (3) both against the committed golden fixtures in tests/golden/
(generated from the compiled reference by tests/golden/gen_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import vectors as V
from tests.conftest import hx

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rows(hexes):
    return np.stack([hx(h) for h in hexes])


def test_x25519_kat(oracles):
    for name, o in oracles.items():
        sk = _rows([k for k, _, _, _ in V.X25519_KAT]); pk = _rows([u for _, u, _, _ in V.X25519_KAT])
        out, skc = o.x25519_shared(pk, sk)
        for i, (_, _, exp, note) in enumerate(V.X25519_KAT):
            assert out[i].tobytes().hex() == exp, (name, note)
        # in-place clamp is part of the contract (curve25519_dh.c:206)
        assert ((skc[:, 0] & 7) == 0).all() and ((skc[:, 31] & 0xC0) == 0x40).all()


def test_x25519_low_order_gives_zero(oracles):
    for name, o in oracles.items():
        pk = _rows(V.X25519_LOW_ORDER_U)
        sk = np.tile(hx(V.X25519_KAT[3][0]), (pk.shape[0], 1))
        out, _ = o.x25519_shared(pk, sk)
        assert not out.any(), name


def test_x25519_public_kat_both_paths(oracles):
    for name, o in oracles.items():
        sk = _rows([k for k, _, _ in V.X25519_PUBLIC_KAT])
        for fast in (True, False):
            pk, _ = o.x25519_public(sk, fast=fast)
            for i, (_, exp, note) in enumerate(V.X25519_PUBLIC_KAT):
                assert pk[i].tobytes().hex() == exp, (name, fast, note)


def test_x25519_iterated(oracles):
    o = oracles["port"]
    k = hx("09" + "00" * 31)[None, :]; u = k.copy()
    for i in range(1000):
        out, _ = o.x25519_shared(u, k)
        u, k = k, out
        if i == 0:
            assert k[0].tobytes().hex() == V.X25519_ITER_1
    assert k[0].tobytes().hex() == V.X25519_ITER_1000


def test_dh_test_keys(oracles):
    """The reference's dh_test (test/curve25519_test.c:429-475): both sides derive the same secret."""
    for name, o in oracles.items():
        a = hx(V.DH_TEST["alice_sk"])[None, :]; b = hx(V.DH_TEST["bruce_sk"])[None, :]
        apk, a_c = o.x25519_public(a); bpk, b_c = o.x25519_public(b)
        assert apk[0].tobytes().hex() == V.DH_TEST["alice_pk"] and bpk[0].tobytes().hex() == V.DH_TEST["bruce_pk"]
        s1, _ = o.x25519_shared(bpk, a); s2, _ = o.x25519_shared(apk, b)
        assert s1[0].tobytes().hex() == V.DH_TEST["shared"] == s2[0].tobytes().hex(), name


def test_ed25519_kat(oracles):
    for name, o in oracles.items():
        for seed, pk, msg, sig in V.ED25519_KAT:
            pub, priv = o.ed25519_keypair(hx(seed)[None, :])
            assert pub[0].tobytes().hex() == pk and priv[0].tobytes().hex() == seed + pk, name
            m = hx(msg); off = np.array([0, m.size], np.uint64)
            s = o.ed25519_sign(priv, m, off)
            assert s[0].tobytes().hex() == sig, name
            assert o.ed25519_verify(s, pub, m, off)[0] == 1
            bad = s.copy(); bad[0, 3] ^= 1
            assert o.ed25519_verify(bad, pub, m, off)[0] == 0
            bad = s.copy(); bad[0, 40] ^= 1
            assert o.ed25519_verify(bad, pub, m, off)[0] == 0


def test_verify_accepts_s_plus_l(oracles):
    """Permissive verification: S + L (no carry out of 256 bits) is accepted (ed25519_verify.c:308)."""
    seed, pk, msg, sig = V.ED25519_KAT[0]
    s = int.from_bytes(bytes.fromhex(sig)[32:], "little") + V.L_ORDER
    assert s < 2**256
    forged = np.frombuffer(bytes.fromhex(sig)[:32] + s.to_bytes(32, "little"), np.uint8)[None, :]
    for name, o in oracles.items():
        assert o.ed25519_verify(forged, hx(pk)[None, :], np.zeros(0, np.uint8), np.array([0, 0], np.uint64))[0] == 1, name


def test_sha512_kat(oracles):
    lib = oracles["port"].lib
    out = (C.c_uint8 * 64)()
    lib.orc_sha512(out, b"abc", C.c_size_t(3))
    assert bytes(out).hex() == V.SHA512_ABC
    # the reference's second SHA-512 KAT: one million 'a' (test/curve25519_selftest.c:131-141)
    m = b"a" * 1000000
    lib.orc_sha512(out, m, C.c_size_t(len(m)))
    assert bytes(out).hex() == ("e718483d0ce769644e2e42c7bc15b4638e1f98b13b2044285632a803afa973eb"
                                "de0ff244877ea60a4cb0432ce577c31beb009c5c2c49aa2e4eadb217ad8cc09b")
    import hashlib
    rng = np.random.Generator(np.random.PCG64(7))
    for n in (0, 1, 55, 111, 112, 113, 127, 128, 129, 239, 240, 241, 1000):
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        lib.orc_sha512(out, m, C.c_size_t(n))
        assert bytes(out) == hashlib.sha512(m).digest(), n


def test_port_primitives_against_python_bigints(oracles, rng):
    lib = oracles["port"].lib
    P, L = V.P_FIELD, V.L_ORDER
    buf = lambda b: (C.c_uint8 * len(b)).from_buffer_copy(b)
    out = (C.c_uint8 * 32)()
    for _ in range(200):
        a = int.from_bytes(rng.bytes(32), "little"); b = int.from_bytes(rng.bytes(32), "little")
        ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")
        lib.orc_fe_mul(out, buf(ab), buf(bb)); assert int.from_bytes(bytes(out), "little") == a * b % P
        lib.orc_fe_add(out, buf(ab), buf(bb)); assert int.from_bytes(bytes(out), "little") == (a + b) % P
        lib.orc_fe_sub(out, buf(ab), buf(bb)); assert int.from_bytes(bytes(out), "little") == (a - b) % P
        lib.orc_sc_reduce64(out, buf(ab + bb)); assert int.from_bytes(bytes(out), "little") == (a + (b << 256)) % L
        c = int.from_bytes(rng.bytes(32), "little")
        lib.orc_sc_muladd(out, buf(ab), buf(bb), buf(c.to_bytes(32, "little")))
        assert int.from_bytes(bytes(out), "little") == (a * b + c) % L
    a = int.from_bytes(rng.bytes(32), "little"); ab = a.to_bytes(32, "little")
    lib.orc_fe_inv(out, buf(ab)); assert int.from_bytes(bytes(out), "little") == pow(a, P - 2, P)
    lib.orc_fe_pow22523(out, buf(ab)); assert int.from_bytes(bytes(out), "little") == pow(a, (P - 5) // 8, P)


def test_port_table_matches_reference_table(oracles):
    """The restatement derives the 8-fold table from first principles; it must equal _w_base_folding8."""
    if "reference" not in oracles:
        pytest.skip("compiled reference not present on this box")
    ref = (C.c_uint8 * (256 * 96)).in_dll(oracles["reference"].lib, "_w_base_folding8")
    ref = np.frombuffer(ref, np.uint8).reshape(256, 96)
    out = (C.c_uint8 * 96)()
    for i in range(256):
        oracles["port"].lib.orc_base_table_entry(out, i)
        assert bytes(out) == ref[i].tobytes(), i


def test_engine_table_matches_reference_table(oracles):
    """The product's generated tables (tools/gen_base_table.py), read from the built library under the reference's own
    symbol names, equal the reference's data: _w_base_folding8 (source/base_folding8.h) entry by entry against the
    restatement's derivation and, when the compiled reference is here, every exported table against its bytes."""
    from curve25519_b200 import build, _native
    build.build()
    L = C.CDLL(_native.LIB_PATH)
    tab = (C.c_uint32 * (256 * 24)).in_dll(L, "_w_base_folding8")
    vals = np.frombuffer(bytes(tab), dtype=np.uint32)
    out = (C.c_uint8 * 96)()
    for i in range(256):
        oracles["port"].lib.orc_base_table_entry(out, i)
        assert bytes(out) == vals[24 * i:24 * i + 24].tobytes(), i
    if "reference" in oracles:
        R = oracles["reference"].lib
        for name, words in (("_w_base_folding8", 256 * 24), ("_w_P", 8), ("_w_2d", 8), ("_w_I", 8), ("_w_NxBPO", 16 * 8)):
            ours = bytes((C.c_uint32 * words).in_dll(L, name)); ref = bytes((C.c_uint32 * words).in_dll(R, name))
            assert ours == ref, name
        assert bytes((C.c_uint8 * 32).in_dll(L, "ecp_BasePoint")) == bytes((C.c_uint8 * 32).in_dll(R, "ecp_BasePoint"))


def test_port_matches_reference_random(oracles, rng):
    if "reference" not in oracles:
        pytest.skip("compiled reference not present on this box")
    R, Pt = oracles["reference"], oracles["port"]
    n = 512
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    a, ask = R.x25519_shared(pk, sk, threads=4); b, bsk = Pt.x25519_shared(pk, sk, threads=4)
    assert (a == b).all() and (ask == bsk).all()
    a, _ = R.x25519_public(sk, fast=True, threads=4); b, _ = Pt.x25519_public(sk, fast=True, threads=4)
    assert (a == b).all()
    pub, priv = R.ed25519_keypair(sk, threads=4); pub2, priv2 = Pt.ed25519_keypair(sk, threads=4)
    assert (pub == pub2).all() and (priv == priv2).all()
    lens = rng.integers(0, 300, n); off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    flat = rng.integers(0, 256, int(off[-1]), dtype=np.uint8)
    sa = R.ed25519_sign(priv, flat, off, threads=4); sb = Pt.ed25519_sign(priv, flat, off, threads=4)
    assert (sa == sb).all()
    bad = sa.copy(); bad[::3, 5] ^= 4; bad[1::7, 40] ^= 1
    va = R.ed25519_verify(bad, pub, flat, off, threads=4); vb = Pt.ed25519_verify(bad, pub, flat, off, threads=4)
    assert (va == vb).all() and 0 < va.sum() < n
    garbage_pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    assert (R.ed25519_verify(sa, garbage_pk, flat, off, threads=4) == Pt.ed25519_verify(sa, garbage_pk, flat, off, threads=4)).all()


@pytest.mark.parametrize("name", ["x25519", "ed25519"])
def test_golden_fixtures(oracles, name):
    """Committed outputs of the compiled reference (tests/golden/gen_golden.py): every checker present must
    reproduce them -- this is what pins the restatement on boxes where the reference itself is absent."""
    g = json.load(open(os.path.join(GOLD, name + ".json")))
    for oname, o in oracles.items():
        if name == "x25519":
            sk = _rows(g["sk"]); pk = _rows(g["pk"])
            out, skc = o.x25519_shared(pk, sk)
            assert [r.tobytes().hex() for r in out] == g["shared"], oname
            assert [r.tobytes().hex() for r in skc] == g["sk_clamped"], oname
            pub, _ = o.x25519_public(sk, fast=True)
            assert [r.tobytes().hex() for r in pub] == g["public"], oname
        else:
            seed = _rows(g["seed"])
            pub, priv = o.ed25519_keypair(seed)
            assert [r.tobytes().hex() for r in pub] == g["pub"], oname
            msgs = [bytes.fromhex(m) for m in g["msg"]]
            off = np.zeros(len(msgs) + 1, np.uint64); off[1:] = np.cumsum([len(m) for m in msgs])
            flat = np.frombuffer(b"".join(msgs), np.uint8)
            sig = o.ed25519_sign(priv, flat, off)
            assert [r.tobytes().hex() for r in sig] == g["sig"], oname
            tampered = _rows(g["sig_tampered"])
            ok = o.ed25519_verify(tampered, pub, flat, off)
            assert ok.tolist() == g["ok_tampered"], oname


def test_asm64_reference_build_agrees_with_portable_c(rng):
    """The optional "best CPU" baseline (the reference's source/asm64 build) computes the same bytes as its portable C."""
    from oracle import pyoracle
    if not (pyoracle.available("reference_asm") and pyoracle.available("reference")):
        pytest.skip("reference builds not present on this box")
    A, R = pyoracle.Oracle("reference_asm"), pyoracle.Oracle("reference")
    n = 256
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    assert (A.x25519_shared(pk, sk, threads=4)[0] == R.x25519_shared(pk, sk, threads=4)[0]).all()
    pa, va = A.ed25519_keypair(sk, threads=4); pr, vr = R.ed25519_keypair(sk, threads=4)
    assert (pa == pr).all()
    sa = A.ed25519_sign(va, msgs, threads=4)
    assert (sa == R.ed25519_sign(vr, msgs, threads=4)).all()
    sa[::3, 5] ^= 1
    assert (A.ed25519_verify(sa, pa, msgs, threads=4) == R.ed25519_verify(sa, pr, msgs, threads=4)).all()
