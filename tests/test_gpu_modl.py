"""GPU: arithmetic modulo the group order and the self-test's split-key / DH known answers
(test/curve25519_selftest.c:101-115, 258-282, 752-817) through the engine."""
import numpy as np
import pytest

from . import vectors as V
from .conftest import hx

pytestmark = pytest.mark.gpu
L = V.L_ORDER


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ints(a):
    return [int.from_bytes(r.tobytes(), "little") for r in a]


def _rows(vals):
    return np.stack([np.frombuffer(int(v).to_bytes(32, "little"), np.uint8) for v in vals])


def test_modl_ops_vs_bigints(engine, rng):
    n = 512
    a = rng.integers(0, 256, (n, 32), dtype=np.uint8); b = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    specials = [0, 1, 2, L - 1, L, L + 1, 15 * L, 2**252, 2**255 - 19, 2**256 - 1]
    for i, v in enumerate(specials):
        a[i] = _rows([v])[0]; b[len(specials) + i] = _rows([v])[0]
    A, B = _ints(a), _ints(b)
    Rinv = pow(2**256, -1, L)
    for op, f in [(engine.MODL_MULMOD, lambda x, y: x * y % L), (engine.MODL_ADDMOD, lambda x, y: (x + y) % L),
                  (engine.MODL_MONTMUL, lambda x, y: x * y * Rinv % L)]:
        got = engine.modl(op, _dev(a), _dev(b)).cpu().numpy()
        assert _ints(got) == [f(x, y) for x, y in zip(A, B)], op
        assert (engine.modl(op, a, b) == got).all()              # host-pointer path
    m = 96                                                          # 256 squarings + 256 multiplications each
    got = engine.modl(engine.MODL_EXPMOD, _dev(a[:m]), _dev(b[:m])).cpu().numpy()
    assert _ints(got) == [pow(x, y, L) for x, y in zip(A[:m], B[:m])]
    got = engine.modl(engine.MODL_INVMOD, _dev(a[:m])).cpu().numpy()
    for x, g in zip(A[:m], _ints(got)):
        assert g == pow(x, L - 2, L)
        if x % L:
            assert g * x % L == 1


def test_selftest_mod_bpo_identities(engine):
    """I*D mod BPO three ways (selftest.c:655-692) and 1/k1 == k2, k1*k2 == 1 (:800-817)."""
    I, D = _rows([V.CONST_I]), _rows([V.CONST_D])
    exp = hx(V.SELFTEST_IXD_MOD_BPO)
    assert (engine.modl(engine.MODL_MULMOD, I, D)[0] == exp).all()
    R2 = _rows([pow(2, 512, L)]); one = _rows([1])
    toM = lambda x: engine.modl(engine.MODL_MONTMUL, x, R2)          # eco_ToMont
    c = engine.modl(engine.MODL_MONTMUL, toM(I), toM(D))
    assert (engine.modl(engine.MODL_MONTMUL, c, one)[0] == exp).all()  # eco_FromMont
    k1, k2 = hx(V.SELFTEST_K1)[None], hx(V.SELFTEST_K2)[None]
    assert (engine.modl(engine.MODL_INVMOD, k1) == k2).all()
    assert (engine.modl(engine.MODL_MULMOD, k1, k2) == one).all()


def test_selftest_dh_and_split_key(engine, oracles):
    """Key generation + ECDH with the self-test's pk1 / pk2 used as RAW scalars through ecp_PointMultiply (:769-784), and
    the split-key round trip k2.(k1.D) == D (:786-798) -- through c25519_x25519_scalarmult_raw_* and the legacy symbol."""
    import ctypes as C
    B = hx("09" + "00" * 31)[None]
    pk1, pk2 = hx(V.SELFTEST_PK1)[None], hx(V.SELFTEST_PK2)[None]
    a = engine.x25519_scalarmult_raw(B, pk1); b = engine.x25519_scalarmult_raw(B, pk2)
    c = engine.x25519_scalarmult_raw(b, pk1); d = engine.x25519_scalarmult_raw(a, pk2)
    assert (c == d).all() and c.any()
    for name, o in oracles.items():                                  # the same four calls on the CPU checkers
        buf = (C.c_uint8 * 32)()
        o.lib.ecp_PointMultiply(buf, B.ctypes.data_as(C.c_void_p), pk1.ctypes.data_as(C.c_void_p), 32)
        assert bytes(buf) == a.tobytes(), name
        o.lib.ecp_PointMultiply(buf, b.ctypes.data_as(C.c_void_p), pk1.ctypes.data_as(C.c_void_p), 32)
        assert bytes(buf) == c.tobytes(), name
    secret = np.full((1, 32), 0x44, np.uint8)
    pub = engine.x25519_scalarmult_raw(B, secret)
    q1 = engine.x25519_scalarmult_raw(pub, hx(V.SELFTEST_K1)[None])
    q0 = engine.x25519_scalarmult_raw(q1, hx(V.SELFTEST_K2)[None])
    assert (q0 == pub).all()
    # batched on device: 1024 random points of the prime-order subgroup, split and re-joined
    import torch
    rng = np.random.Generator(np.random.PCG64(7))
    sk = rng.integers(0, 256, (1024, 32), dtype=np.uint8)
    pts, _ = engine.x25519_public(torch.from_numpy(sk).cuda())
    k1 = torch.from_numpy(np.repeat(hx(V.SELFTEST_K1)[None], 1024, 0)).cuda()
    k2 = torch.from_numpy(np.repeat(hx(V.SELFTEST_K2)[None], 1024, 0)).cuda()
    back = engine.x25519_scalarmult_raw(engine.x25519_scalarmult_raw(pts, k1), k2)
    assert torch.equal(back, pts)


def test_point_multiply_refuses_wide_scalars():
    """ecp_PointMultiply with len > 32: zero high bytes are accepted (same value), non-zero ones are refused loudly."""
    import ctypes as C
    import subprocess
    import sys
    from curve25519_b200 import _native
    Lb = _native.lib()
    q = (C.c_uint8 * 32)(); q2 = (C.c_uint8 * 32)()
    p = (C.c_uint8 * 32)(9)
    k = (C.c_uint8 * 40)(*([5] + [0] * 39))
    Lb.ecp_PointMultiply(q, p, k, 40)
    Lb.ecp_PointMultiply(q2, p, k, 1)
    assert bytes(q) == bytes(q2)
    code = ("import ctypes as C; from curve25519_b200 import _native; L=_native.lib();"
            "q=(C.c_uint8*32)(); p=(C.c_uint8*32)(9); k=(C.c_uint8*40)(*([1]*40)); L.ecp_PointMultiply(q,p,k,40)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "wider than 256 bits" in r.stderr
