"""GPU: the reference's INTERNAL word-level symbols exported by the library (csrc/legacy_internals.cu; prototypes
source/curve25519_mehdi.h:93-160, source/sha512.h:85-87), each against Python integers / hashlib.  The reference's own
self-test drives the same symbols end to end in tests/test_gpu_dropin.py; this file pins them one by one, including the
representation contract (results below 2^256, congruent; canonical for ecp_Mod / ecp_MulMod / eco_*)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from . import vectors as V

pytestmark = pytest.mark.gpu
P, L = V.P_FIELD, V.L_ORDER
W8 = C.c_uint32 * 8


def _w(x):
    return W8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def _i(w):
    return sum(int(v) << (32 * i) for i, v in enumerate(w))


@pytest.fixture(scope="module")
def lib(engine):
    from curve25519_b200 import _native
    Lb = C.CDLL(_native.LIB_PATH)
    Lb.ecp_Add.restype = C.c_uint32
    Lb.ecp_Sub.restype = C.c_int32
    Lb.ecp_CmpNE.restype = C.c_int
    Lb.ecp_DecodeInt.restype = C.c_ubyte
    return Lb


def _vals(rng, n):
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 1, 2**256 - 39, 2**256 - 38, 2**256 - 1, 2**255, L - 1, L, L + 1, 15 * L]
    return edge + [int.from_bytes(rng.bytes(32), "little") for _ in range(n)]


def test_field_words(lib, rng):
    vals = _vals(rng, 40)
    z = W8(); t16 = (C.c_uint32 * 16)()
    for k, a in enumerate(vals):
        b = vals[(7 * k + 3) % len(vals)]
        lib.ecp_MulReduce(z, _w(a), _w(b)); assert _i(z) < 2**256 and _i(z) % P == a * b % P
        lib.ecp_SqrReduce(z, _w(a)); assert _i(z) % P == a * a % P
        lib.ecp_AddReduce(z, _w(a), _w(b)); assert _i(z) % P == (a + b) % P
        lib.ecp_SubReduce(z, _w(a), _w(b)); assert _i(z) % P == (a - b) % P
        lib.ecp_MulMod(z, _w(a), _w(b)); assert _i(z) == a * b % P
        x = _w(a); lib.ecp_Mod(x); assert _i(x) == a % P
        lib.ecp_Mul(t16, _w(a), _w(b)); assert sum(int(v) << (32 * i) for i, v in enumerate(t16)) == a * b
        c = lib.ecp_Add(z, _w(a), _w(b)); assert _i(z) + (c << 256) == a + b
        bw = lib.ecp_Sub(z, _w(a), _w(b)); assert bw in (0, -1) and _i(z) == (a - b) % 2**256 and (bw == -1) == (a < b)
    for a in vals[:12]:
        lib.ecp_Inverse(z, _w(a)); assert _i(z) % P == pow(a, P - 2, P)
    # aliasing like the reference's callers use it (Z == X)
    x = _w(vals[20]); lib.ecp_MulReduce(x, x, x); assert _i(x) % P == vals[20] ** 2 % P


def test_order_words(lib, rng):
    vals = _vals(rng, 40)
    z = W8()
    for k, a in enumerate(vals):
        b = vals[(5 * k + 1) % len(vals)]
        lib.eco_MulReduce(z, _w(a), _w(b)); assert _i(z) == a * b % L
        lib.eco_AddReduce(z, _w(a), _w(b)); assert _i(z) == (a + b) % L
        x = _w(a); lib.eco_Mod(x); assert _i(x) == a % L
        hi = (k * 0x9E3779B1) & 0xFFFFFFFF
        lib.eco_ReduceHiWord(z, C.c_uint32(hi), _w(a)); assert _i(z) == (a + (hi << 256)) % L
        md = rng.bytes(64)
        lib.eco_DigestToWords(z, md); assert _i(z) == int.from_bytes(md, "little") % L


def test_codecs_and_tables(lib):
    y = W8(); buf = (C.c_ubyte * 32)()
    raw = bytes(range(1, 32)) + bytes([0x85])
    assert lib.ecp_DecodeInt(y, raw) == 1 and _i(y) == int.from_bytes(raw, "little") & (2**255 - 1)
    lib.ecp_EncodeInt(buf, y, C.c_ubyte(1)); assert bytes(buf) == raw
    lib.ecp_BytesToWords(y, raw); assert _i(y) == int.from_bytes(raw, "little")
    lib.ecp_WordsToBytes(buf, y); assert bytes(buf) == raw
    assert lib.ecp_CmpNE(_w(5), _w(5)) == 0 and lib.ecp_CmpNE(_w(5), _w(6)) != 0
    assert _i(W8.in_dll(lib, "_w_P")) == P
    nx = ((C.c_uint32 * 8) * 16).in_dll(lib, "_w_NxBPO")
    assert [_i(r) for r in nx] == [i * L for i in range(16)]
    assert pow(_i(W8.in_dll(lib, "_w_I")), 2, P) == P - 1


def test_edwards_point_words(lib, oracles):
    """edp_BasePointMultiply / edp_DoublePoint / edp_AddBasePoint / edp_AddPoint / ed25519_UnpackPoint: k*B through the comb
    equals k*B through double-and-add on the exported primitives, and both equal the public key the checkers derive."""
    d = (-121665 * pow(121666, P - 2, P)) % P

    def on_curve(x, y):
        return (-x * x + y * y - 1 - d * x * x * y * y) % P == 0
    aff = (C.c_uint32 * 16)()
    for k in (1, 2, 3, 7, 127, 2**200 + 12345, L - 1):
        lib.edp_BasePointMultiply(aff, _w(k), None)
        x, y = _i(aff[:8]), _i(aff[8:])
        assert x < P and y < P and on_curve(x, y), k
        # the same multiple by double-and-add over Ext_POINT words
        ext = (C.c_uint32 * 32)(); ext[8] = 1; ext[16] = 1                 # neutral element (0, 1, 1, 0)
        for bit in bin(k)[2:]:
            lib.edp_DoublePoint(ext)
            if bit == "1":
                lib.edp_AddBasePoint(ext)
        X, Y, Z = _i(ext[:8]), _i(ext[8:16]), _i(ext[16:24])
        zi = pow(Z, P - 2, P)
        assert (X * zi % P, Y * zi % P) == (x, y), k
        enc = (y | ((x & 1) << 255)).to_bytes(32, "little")
        back = (C.c_uint32 * 16)()
        lib.ed25519_UnpackPoint(back, enc)
        assert (_i(back[:8]) % P, _i(back[8:])) == (x, y)
    # edp_AddPoint with a PE operand built from an affine point: (x, y) + (x, y) == 2 (x, y)
    lib.edp_BasePointMultiply(aff, _w(5), None)
    x, y = _i(aff[:8]), _i(aff[8:])
    pe = (C.c_uint32 * 32)(*(list(_w((y + x) % P)) + list(_w((y - x) % P)) + list(_w(2 * d * x * y % P)) + list(_w(2))))
    ext = (C.c_uint32 * 32)(*(list(_w(x)) + list(_w(y)) + list(_w(1)) + list(_w(x * y % P))))
    r = (C.c_uint32 * 32)()
    lib.edp_AddPoint(r, ext, pe)
    zi = pow(_i(r[16:24]), P - 2, P)
    lib.edp_BasePointMultiply(aff, _w(10), None)
    assert (_i(r[:8]) * zi % P, _i(r[8:16]) * zi % P) == (_i(aff[:8]), _i(aff[8:]))


def test_streaming_sha512(lib, rng):
    ctx = (C.c_ubyte * 216)(); md = (C.c_ubyte * 64)()
    for total, pieces in [(0, [0]), (3, [3]), (111, [50, 61]), (112, [112]), (128, [1, 127]), (129, [128, 1]), (1000, [7, 300, 693]),
                          (70000, [65536, 4464]), (200, [0, 100, 0, 100])]:
        msg = rng.bytes(total)
        lib.SHA512_Init(ctx)
        o = 0
        for p in pieces:
            lib.SHA512_Update(ctx, msg[o:o + p], C.c_size_t(p)); o += p
        lib.SHA512_Final(md, ctx)
        assert bytes(md) == hashlib.sha512(msg).digest(), (total, pieces)
