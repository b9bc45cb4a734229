"""CPU-only: the C-ABI library loads and exports every symbol the public headers declare, and every
compute entry point fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:c25519|curve25519_dh|ed25519|ecp|eco|edp|SHA512)_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from curve25519_b200 import build, _native
    build.build()
    L = C.CDLL(_native.LIB_PATH)
    names = _declared("c25519_b200.h") + _declared("c25519_legacy.h")
    assert len(names) >= 75
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    _native.lib()          # also checks the ctypes prototypes bind


def test_legacy_header_names_forward():
    for h in ("curve25519_dh.h", "ed25519_signature.h"):
        assert "c25519_legacy.h" in open(os.path.join(ROOT, "include", h)).read()


def test_compute_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; this test is for GPU-less boxes")
    except ImportError:
        pass
    from curve25519_b200 import api, EngineError
    z = np.zeros((2, 32), np.uint8)
    with pytest.raises(EngineError, match="no CUDA device"):
        api.x25519_shared(z, z)
    with pytest.raises(EngineError):
        api.ed25519_keypair(z)


def test_product_does_not_reference_the_oracle():
    """Nothing under curve25519_b200/ may import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "curve25519_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt and "pyoracle" not in txt and "liboracle" not in txt and "libref25519" not in txt, f


def test_host_side_argument_validation_needs_no_gpu():
    """api.py rejects malformed arrays before anything reaches native code: caller-supplied outputs must be
    exactly what the C ABI writes; record arrays must have the record width of the operation."""
    from curve25519_b200 import api
    z = np.zeros((4, 32), np.uint8)
    for bad in (np.zeros((4, 16), np.uint8), np.zeros((8, 32), np.uint8)[::2], np.zeros((4, 32), np.int32), np.zeros((3, 32), np.uint8)):
        with pytest.raises(ValueError):
            api.x25519_shared(z, z, out=bad)
        with pytest.raises(ValueError):
            api.x25519_public(z, out=bad)
    with pytest.raises(ValueError):
        api.x25519_shared(np.zeros((4, 31), np.uint8), z)
    with pytest.raises(ValueError):
        api.x25519_shared(np.zeros((5, 32), np.uint8), z)
    with pytest.raises(ValueError):
        api.ed25519_sign(np.zeros((4, 64), np.uint8), np.zeros(10, np.uint8), np.zeros(4, np.uint64))      # needs n + 1 offsets
    with pytest.raises(ValueError):
        api.ed25519_verify(np.zeros((4, 64), np.uint8), np.zeros((4, 32), np.uint8), np.zeros((3, 8), np.uint8))


def test_sharded_entry_points_fail_loudly_without_gpu_or_nccl():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; this test is for GPU-less boxes")
    except ImportError:
        pass
    from curve25519_b200 import _native
    L = _native.lib()
    buf = (C.c_uint8 * 128)()
    rc = L.c25519_nccl_unique_id(buf)
    assert rc != 0 or any(buf)                       # either NCCL is absent (error, message set) or it answered
    assert L.c25519_x25519_shared_sharded(None, None, None, 4, None, None) != 0
    assert L.c25519_last_error()
