"""GPU parity for the X25519 path, through the C ABI (device-pointer and host-pointer flavours and the
n = 1 legacy wrappers), against the CPU checkers, the RFC 7748 vectors and the committed golden fixtures.
Bit-exact: outputs AND the in-place-clamped secret keys must match byte for byte."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import vectors as V
from tests.conftest import hx

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rows(hexes):
    return np.stack([hx(h) for h in hexes])


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_kat_vectors(engine):
    sk = _rows([k for k, _, _, _ in V.X25519_KAT]); pk = _rows([u for _, u, _, _ in V.X25519_KAT])
    out, skc = engine.x25519_shared(_dev(pk), _dev(sk))
    out = out.cpu().numpy(); skc = skc.cpu().numpy()
    for i, (_, _, exp, note) in enumerate(V.X25519_KAT):
        assert out[i].tobytes().hex() == exp, note
    exp_sk = sk.copy(); exp_sk[:, 0] &= 0xF8; exp_sk[:, 31] = (exp_sk[:, 31] | 0x40) & 0x7F
    assert (skc == exp_sk).all()


def test_low_order_points_give_zero(engine):
    pk = _rows(V.X25519_LOW_ORDER_U)
    sk = np.tile(hx(V.X25519_KAT[3][0]), (pk.shape[0], 1))
    out, _ = engine.x25519_shared(_dev(pk), _dev(sk))
    assert not out.cpu().numpy().any()


def test_golden_fixture(engine):
    g = json.load(open(os.path.join(GOLD, "x25519.json")))
    sk = _rows(g["sk"]); pk = _rows(g["pk"])
    out, skc = engine.x25519_shared(_dev(pk), _dev(sk))
    assert [r.tobytes().hex() for r in out.cpu().numpy()] == g["shared"]
    assert [r.tobytes().hex() for r in skc.cpu().numpy()] == g["sk_clamped"]
    for ladder in (True, False):
        pub, _ = engine.x25519_public(_dev(sk), ladder=ladder)
        assert [r.tobytes().hex() for r in pub.cpu().numpy()] == g["public"], ladder


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 1000, 20000])
def test_random_batches_vs_oracle(engine, oracle, rng, n):
    """Ragged batch sizes around warp / CTA boundaries; uniform random scalars and points, no pre-clamping,
    no bit-255 masking."""
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    exp, exp_sk = oracle.x25519_shared(pk, sk, threads=os.cpu_count() or 1)
    out, skc = engine.x25519_shared(_dev(pk), _dev(sk))
    assert (out.cpu().numpy() == exp).all() and (skc.cpu().numpy() == exp_sk).all()
    # host-pointer flavour (H2D + kernel + D2H inside the call)
    out_h, skc_h = engine.x25519_shared(pk, sk)
    assert (out_h == exp).all() and (skc_h == exp_sk).all()


def test_empty_batch(engine):
    z = np.zeros((0, 32), np.uint8)
    out, skc = engine.x25519_shared(z, z)
    assert out.shape == (0, 32)
    out, skc = engine.x25519_shared(_dev(z), _dev(z))
    assert tuple(out.shape) == (0, 32)


def test_public_ladder_vs_oracle(engine, oracle, rng):
    sk = rng.integers(0, 256, (3000, 32), dtype=np.uint8)
    exp, exp_sk = oracle.x25519_public(sk, fast=True, threads=os.cpu_count() or 1)
    for ladder in (True, False):
        out, skc = engine.x25519_public(_dev(sk), ladder=ladder)
        assert (out.cpu().numpy() == exp).all() and (skc.cpu().numpy() == exp_sk).all(), ladder
    out, skc = engine.x25519_public(sk, ladder=False)
    assert (out == exp).all() and (skc == exp_sk).all()


def test_iterated_rfc7748(engine):
    """RFC 7748 5.2 iterated test, 1 and 1000 iterations, through the n = 1 host path."""
    k = hx("09" + "00" * 31)[None, :]; u = k.copy()
    for i in range(1000):
        out, _ = engine.x25519_shared(u, k)
        u, k = k, out
        if i == 0:
            assert k[0].tobytes().hex() == V.X25519_ITER_1
    assert k[0].tobytes().hex() == V.X25519_ITER_1000


def test_legacy_wrappers(engine):
    """The reference's own dh_test (test/curve25519_test.c:429-475) through the re-exported legacy symbols."""
    from curve25519_b200 import _native
    L = _native.lib()
    a = (C.c_uint8 * 32).from_buffer_copy(bytes.fromhex(V.DH_TEST["alice_sk"]))
    b = (C.c_uint8 * 32).from_buffer_copy(bytes.fromhex(V.DH_TEST["bruce_sk"]))
    apk = (C.c_uint8 * 32)(); bpk = (C.c_uint8 * 32)(); s1 = (C.c_uint8 * 32)(); s2 = (C.c_uint8 * 32)()
    L.curve25519_dh_CalculatePublicKey(apk, a)
    L.curve25519_dh_CalculatePublicKey_fast(bpk, b)
    assert bytes(apk).hex() == V.DH_TEST["alice_pk"] and bytes(bpk).hex() == V.DH_TEST["bruce_pk"]
    assert (a[0] & 7) == 0 and (a[31] & 0xC0) == 0x40          # clamped in place
    L.curve25519_dh_CreateSharedKey(s1, bpk, a)
    L.curve25519_dh_CreateSharedKey(s2, apk, b)
    assert bytes(s1).hex() == V.DH_TEST["shared"] == bytes(s2).hex()


def test_full_size_1m_properties(engine, oracle, rng):
    """BASELINE config 2 at full size (2^20 ops): size-independent properties plus a sampled oracle check.
      * Diffie-Hellman commutativity: shared(a, pub(b)) == shared(b, pub(a)) for every pair
      * 4096 randomly chosen records equal the oracle's output byte for byte
      * clamping is idempotent and exactly the RFC mask"""
    import torch
    n = 1 << 20
    sk_a = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk_b = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    da, db = _dev(sk_a), _dev(sk_b)
    pa, ca = engine.x25519_public(da, ladder=False); pb, cb = engine.x25519_public(db, ladder=False)
    s1, ca2 = engine.x25519_shared(pb, da); s2, _ = engine.x25519_shared(pa, db)
    assert torch.equal(s1, s2)
    assert torch.equal(ca, ca2)
    exp_sk = sk_a.copy(); exp_sk[:, 0] &= 0xF8; exp_sk[:, 31] = (exp_sk[:, 31] | 0x40) & 0x7F
    assert (ca.cpu().numpy() == exp_sk).all()
    # SURVEY 8d config 2 asks for 100 % byte equality on the whole batch: with >= 32 host cores the oracle does all
    # 2^20 records in seconds; on smaller hosts a random 4096-record sample is checked instead.
    ncpu = os.cpu_count() or 1
    idx = np.arange(n) if ncpu >= 32 else rng.choice(n, 4096, replace=False)
    pb_h = pb.cpu().numpy()
    exp, _ = oracle.x25519_shared(pb_h[idx], sk_a[idx], threads=ncpu)
    assert (s1.cpu().numpy()[idx] == exp).all()
    # and uniformly random (mostly off-curve / twist, bit 255 set in half) peer points at full size
    pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    s3, c3 = engine.x25519_shared(_dev(pk), da)
    exp, exp_sk = oracle.x25519_shared(pk[idx], sk_a[idx], threads=ncpu)
    assert (s3.cpu().numpy()[idx] == exp).all() and (c3.cpu().numpy()[idx] == exp_sk).all()


def test_generic_scalarmult_unclamped(engine, oracles, rng):
    """ecp_PointMultiply (curve25519_dh.c:94-157) with ARBITRARY scalars: no clamping, leading zero bits, K = 0 -> zeros,
    bit 255 of K honoured -- against the checkers' own ecp_PointMultiply, record by record; plus the reference
    self-test's group-order identities through the ladder ((L-1) B and (L+1) B share B's u-coordinate, L B -> zeros;
    test/curve25519_selftest.c:752-767)."""
    n = 600
    k = rng.integers(0, 256, (n, 32), dtype=np.uint8); u = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    L = V.L_ORDER
    specials = [0, 1, 2, 3, 8, L - 1, L, L + 1, 2 * L, 2**255 - 19, 2**255, 2**256 - 1, 2**254, 1 << 200, 0x8000000000000000]
    for i, s in enumerate(specials):
        k[i] = np.frombuffer(s.to_bytes(32, "little"), np.uint8)
    for i, h in enumerate(V.X25519_LOW_ORDER_U):
        u[40 + i] = hx(h)
    u[:20] = hx("09" + "00" * 31)
    out = engine.x25519_scalarmult_raw(_dev(u), _dev(k)).cpu().numpy()
    out_h = engine.x25519_scalarmult_raw(u, k)
    assert (out == out_h).all()
    for name, o in oracles.items():
        buf = (C.c_uint8 * 32)()
        for i in range(n):
            o.lib.ecp_PointMultiply(buf, u[i].ctypes.data_as(C.c_void_p), k[i].ctypes.data_as(C.c_void_p), 32)
            assert bytes(buf) == out[i].tobytes(), (name, i, k[i].tobytes().hex(), u[i].tobytes().hex())
    nine = "09" + "00" * 31
    assert out[5].tobytes().hex() == nine and out[7].tobytes().hex() == nine      # (L-1) B, (L+1) B
    assert not out[0].any() and not out[6].any() and not out[8].any()             # 0 B, L B, 2L B
    # the legacy symbol, with a short scalar (len < 32)
    from curve25519_b200 import _native
    q = (C.c_uint8 * 32)(); kk = (C.c_uint8 * 2)(0x39, 0x05)
    _native.lib().ecp_PointMultiply(q, u[0].ctypes.data_as(C.c_void_p), kk, 2)
    for name, o in oracles.items():
        buf = (C.c_uint8 * 32)()
        o.lib.ecp_PointMultiply(buf, u[0].ctypes.data_as(C.c_void_p), kk, 2)
        assert bytes(buf) == bytes(q), name


@pytest.mark.parametrize("key_size", [1, 16, 32, 48, 64])
def test_shared_key_kdf_matches_cxx_wrapper(engine, oracle, rng, key_size):
    """X25519Private::CreateSharedKey (C++/x25519.cpp:75-95): SHA-512 of the shared secret, truncated."""
    import hashlib
    n = 700
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    secret, _ = oracle.x25519_shared(pk, sk, threads=os.cpu_count() or 1)
    got = engine.x25519_shared_kdf(_dev(pk), _dev(sk), key_size).cpu().numpy()
    for i in range(n):
        assert got[i].tobytes() == hashlib.sha512(secret[i].tobytes()).digest()[:key_size], i


def test_abi_argument_checks(engine):
    """Error behaviour of the batched ABI: misaligned device record arrays, null pointers and bad key sizes are refused
    with C25519_E_BAD_ARGUMENT (-3) and a message; nothing is launched."""
    import torch
    from curve25519_b200 import _native
    L = _native.lib()
    buf = torch.zeros(4 * 32 + 64, dtype=torch.uint8, device="cuda")
    base = buf.data_ptr()
    base += (-base) % 32
    before = L.c25519_launch_count()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.c25519_x25519_shared_batch(C.c_void_p(base + 1), C.c_void_p(base + 32), C.c_void_p(base + 64), 1, s) == -3
    assert b"aligned" in L.c25519_last_error()
    assert L.c25519_x25519_shared_batch(None, C.c_void_p(base + 32), C.c_void_p(base + 64), 1, s) == -3
    assert L.c25519_x25519_shared_kdf_batch(C.c_void_p(base), 0, C.c_void_p(base + 32), C.c_void_p(base + 64), 1, s) == -3
    assert L.c25519_x25519_shared_kdf_batch(C.c_void_p(base), 65, C.c_void_p(base + 32), C.c_void_p(base + 64), 1, s) == -3
    assert L.c25519_launch_count() == before
    # n = 0 is a no-op that succeeds even with null pointers
    assert L.c25519_x25519_shared_batch(None, None, None, 0, s) == 0


def test_one_operation_per_warp_kernel_edge_operands(engine, oracles, rng):
    """n <= 592 runs k_x25519_ladder_warp (limb per lane, cross-lane carry look-ahead).  Operands chosen to exercise the carry
    machinery: all-ones limbs, values around p, 2p and 2^256 - 38, limbs of 0xffffffff next to small ones, low-order points."""
    P = V.P_FIELD
    vals = [0, 1, 2, 9, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 1, 2**256 - 39, 2**256 - 38, 2**256 - 37, 2**256 - 1, 2**255, 2**255 - 1,
            2**255 + 18, 2**255 + 19, (1 << 224) - 1, ((1 << 256) - 1) ^ ((1 << 32) - 1), 0xffffffff, 0xffffffff << 32, (1 << 256) - (1 << 32),
            int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16), int("ffffffda" + "ffffffff" * 7, 16)]
    n = 592
    u = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for i, v in enumerate(vals):
        u[i] = np.frombuffer(int(v).to_bytes(32, "little"), np.uint8)
    for i, h in enumerate(V.X25519_LOW_ORDER_U):
        u[40 + i] = hx(h)
    sk[60] = 0xFF; sk[61] = 0; sk[62, :] = 0; sk[62, 0] = 8
    for name, o in oracles.items():
        exp, exp_sk = o.x25519_shared(u, sk, threads=os.cpu_count() or 1)
        for m in (n, 300, 37, 1):                       # several launch shapes of the warp kernel
            out, skc = engine.x25519_shared(_dev(u[:m]), _dev(sk[:m]))
            assert (out.cpu().numpy() == exp[:m]).all() and (skc.cpu().numpy() == exp_sk[:m]).all(), (name, m)
        pub, _ = engine.x25519_public(_dev(sk), ladder=True)
        assert (pub.cpu().numpy() == o.x25519_public(sk, fast=False, threads=os.cpu_count() or 1)[0]).all(), name
