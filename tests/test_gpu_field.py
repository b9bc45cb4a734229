"""GPU: every device primitive, one operation per thread through c25519_test_primitive, against Python
big-integer arithmetic and the CPU checkers' primitives (differential fuzz).
Bit-exact: results are compared after canonicalisation (fe_canon == ecp_Mod)."""
import ctypes as C

import numpy as np
import pytest

from tests import vectors as V

pytestmark = pytest.mark.gpu
P = V.P_FIELD


def _ints(a):
    return [int.from_bytes(r.tobytes(), "little") for r in a]


def _edge_rows():
    vals = [0, 1, 2, 19, 37, 38, 39, P - 1, P, P + 1, 2 * P, 2 * P + 1, 2**255 - 1, 2**255, 2**255 + 18, 2**255 + 19,
            2**256 - 1, 2**256 - 38, 2**256 - 37, 2**256 - 39, 2**32 - 1, 2**32, 2**224, (2**256 - 1) ^ (2**128 - 1), 2**128 - 1,
            0x5555555555555555555555555555555555555555555555555555555555555555, 0xAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA]
    return np.stack([np.frombuffer(v.to_bytes(32, "little"), np.uint8) for v in vals])


def _inputs(rng, n):
    e = _edge_rows()
    m = e.shape[0]
    a = np.concatenate([np.repeat(e, m, axis=0), rng.integers(0, 256, (n, 32), dtype=np.uint8)])
    b = np.concatenate([np.tile(e, (m, 1)), rng.integers(0, 256, (n, 32), dtype=np.uint8)])
    return a, b


@pytest.mark.parametrize("op,fn", [
    (0, lambda a, b: a * b % P), (1, lambda a, b: a * a % P), (2, lambda a, b: (a + b) % P), (3, lambda a, b: (a - b) % P),
    (6, lambda a, b: (a + 121665 * b) % P), (7, lambda a, b: a % P), (10, lambda a, b: (a * a + b * b) % P)])
def test_field_ops_vs_bigint(engine, rng, op, fn):
    import torch
    a, b = _inputs(rng, 4096)
    out = engine.test_primitive(op, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    exp = [fn(x, y) for x, y in zip(_ints(a), _ints(b))]
    got = _ints(out)
    bad = [i for i in range(len(exp)) if exp[i] != got[i]]
    assert not bad, "op %d: %d mismatches, first at %d: a=%x b=%x got=%x exp=%x" % (
        op, len(bad), bad[0], _ints(a)[bad[0]], _ints(b)[bad[0]], got[bad[0]], exp[bad[0]])


@pytest.mark.parametrize("op,e", [(4, P - 2), (5, (P - 5) // 8)])
def test_field_powers_vs_bigint(engine, rng, op, e):
    import torch
    a, _ = _inputs(rng, 512)
    a = np.concatenate([_edge_rows(), a[-512:]])
    out = engine.test_primitive(op, torch.from_numpy(a).cuda()).cpu().numpy()
    assert _ints(out) == [pow(x, e, P) for x in _ints(a)]


def test_field_mul_vs_oracle_primitive(engine, oracles, rng):
    """Same inputs through the checker's own field multiply (the restatement's fe_mul; and the reference's
    ecp_MulReduce + ecp_Mod when the compiled reference is here)."""
    import torch
    a, b = _inputs(rng, 2048)
    out = engine.test_primitive(0, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    lib = oracles["port"].lib
    buf = (C.c_uint8 * 32)()
    for i in range(0, a.shape[0], 7):
        lib.orc_fe_mul(buf, a[i].ctypes.data_as(C.c_void_p), b[i].ctypes.data_as(C.c_void_p))
        assert bytes(buf) == out[i].tobytes()
    if "reference" in oracles:
        R = oracles["reference"].lib
        z = (C.c_uint32 * 8)()
        for i in range(0, a.shape[0], 7):
            R.ecp_MulReduce(z, a[i].ctypes.data_as(C.c_void_p), b[i].ctypes.data_as(C.c_void_p))
            R.ecp_Mod(z)
            assert bytes(z) == out[i].tobytes()


def test_comb_index_extraction(engine, oracles, rng):
    """ecp_8Folds / ecp_4Folds (curve25519_utils.c:144 / :125): the 32 / 64 comb indices of a scalar, against the
    restatement's folds and the reference's exported functions."""
    import torch
    n = 512
    a = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    a[0] = 0; a[1] = 255; a[2] = np.arange(32, dtype=np.uint8)
    f8 = engine.test_primitive(12, torch.from_numpy(a).cuda(), out_rec=32).cpu().numpy()
    f4 = engine.test_primitive(13, torch.from_numpy(a).cuda(), out_rec=64).cpu().numpy()
    b8 = (C.c_uint8 * 32)(); b4 = (C.c_uint8 * 64)()
    for i in range(n):
        oracles["port"].lib.orc_folds8(b8, a[i].ctypes.data_as(C.c_void_p)); assert bytes(b8) == f8[i].tobytes(), i
        oracles["port"].lib.orc_folds4(b4, a[i].ctypes.data_as(C.c_void_p)); assert bytes(b4) == f4[i].tobytes(), i
        if "reference" in oracles:
            oracles["reference"].lib.ecp_8Folds(b8, a[i].ctypes.data_as(C.c_void_p)); assert bytes(b8) == f8[i].tobytes(), i
            oracles["reference"].lib.ecp_4Folds(b4, a[i].ctypes.data_as(C.c_void_p)); assert bytes(b4) == f4[i].tobytes(), i
