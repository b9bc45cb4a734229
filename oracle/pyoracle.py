"""ctypes front-end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

    Oracle("port")       -> oracle/liboracle25519.so   (our C restatement, oracle25519.c)
    Oracle("reference")  -> oracle/_ref/libref25519.so (the reference's own portable-C sources,
                                                       compiled unmodified by oracle/Makefile)
    Oracle("reference_asm") -> oracle/_ref/libref25519_asm.so (the reference's asm64 build, "best CPU" baseline)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  Nothing under curve25519_b200/ does.

All batch methods take/return numpy uint8 arrays with one fixed-size record per row and drive the
library's n=1 legacy API (include/curve25519_dh.h:34-48, include/ed25519_signature.h:40-93 of the
reference) from `threads` worker threads via oracle/batch_driver.c.  Each call also records the
wall-clock seconds of its parallel region in `.last_seconds`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {
    "port": os.path.join(_HERE, "liboracle25519.so"),
    "reference": os.path.join(_HERE, "_ref", "libref25519.so"),
    # the reference's x86-64 assembly build (source/asm64), its fastest CPU configuration: optional second baseline
    "reference_asm": os.path.join(_HERE, "_ref", "libref25519_asm.so"),
}


def build(quiet=True):
    """Compile oracle25519.c and, when /root/reference is present, the reference itself."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def available(kind):
    return os.path.exists(PATHS[kind])


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, kind="port"):
        if kind not in PATHS:
            raise ValueError(kind)
        if not os.path.exists(PATHS[kind]):
            if kind == "port":
                build()
            else:
                raise FileNotFoundError(PATHS[kind] + " (build it in the authoring container: make -C oracle ref)")
        self.kind = kind
        self.lib = C.CDLL(PATHS[kind])
        L = self.lib
        dbl = C.c_double
        L.drv_x25519_shared.restype = dbl
        L.drv_x25519_shared.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.drv_x25519_public.restype = dbl
        L.drv_x25519_public.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.drv_ed25519_keypair.restype = dbl
        L.drv_ed25519_keypair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.drv_ed25519_sign.restype = dbl
        L.drv_ed25519_sign.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
        L.drv_ed25519_verify.restype = dbl
        L.drv_ed25519_verify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_size_t, C.c_int]
        L.ed25519_Verify_Init.restype = C.c_void_p
        L.ed25519_Verify_Init.argtypes = [C.c_void_p, C.c_void_p]
        L.ed25519_Verify_Check.restype = C.c_int
        L.ed25519_Verify_Check.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ed25519_Verify_Finish.argtypes = [C.c_void_p]
        L.ed25519_VerifySignature.restype = C.c_int
        L.ed25519_VerifySignature.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ed25519_SignMessage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ed25519_CreateKeyPair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ed25519_Blinding_Init.restype = C.c_void_p
        L.ed25519_Blinding_Init.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ed25519_Blinding_Finish.argtypes = [C.c_void_p]
        self.last_seconds = 0.0

    # ---------------------------------------------------------------- X25519
    def x25519_shared(self, pk, sk, threads=1):
        """-> (shared[n,32], clamped_sk[n,32]); curve25519_dh_CreateSharedKey, dh.c:201."""
        pk = _u8(pk); sk = _u8(sk).copy(); n = sk.shape[0]
        out = np.zeros((n, 32), np.uint8)
        self.last_seconds = self.lib.drv_x25519_shared(_p(out), _p(pk), _p(sk), n, threads)
        return out, sk

    def x25519_public(self, sk, fast=True, threads=1):
        """-> (pk[n,32], clamped_sk[n,32]); curve25519_dh_CalculatePublicKey[_fast], dh.c:182/192."""
        sk = _u8(sk).copy(); n = sk.shape[0]
        out = np.zeros((n, 32), np.uint8)
        self.last_seconds = self.lib.drv_x25519_public(_p(out), _p(sk), n, 1 if fast else 0, threads)
        return out, sk

    # ---------------------------------------------------------------- Ed25519
    def ed25519_keypair(self, seed, threads=1):
        """-> (pub[n,32], priv[n,64]); ed25519_CreateKeyPair, ed25519_sign.c:344."""
        seed = _u8(seed); n = seed.shape[0]
        pub = np.zeros((n, 32), np.uint8); priv = np.zeros((n, 64), np.uint8)
        self.last_seconds = self.lib.drv_ed25519_keypair(_p(pub), _p(priv), _p(seed), n, threads)
        return pub, priv

    @staticmethod
    def _msgs(msgs, off, n):
        if off is None:
            msgs = _u8(msgs)
            fixed = msgs.shape[1] if msgs.ndim == 2 else 0
            if msgs.size == 0:
                msgs = np.zeros(1, np.uint8)
            return msgs, None, fixed
        off = np.ascontiguousarray(off, dtype=np.uint64)
        assert off.shape[0] == n + 1
        msgs = _u8(msgs)
        if msgs.size == 0:
            msgs = np.zeros(1, np.uint8)
        return msgs, off, 0

    def ed25519_sign(self, priv, msgs, off=None, threads=1):
        """-> sig[n,64]; ed25519_SignMessage, ed25519_sign.c:372.  msgs: [n,len] or flat bytes + off[n+1]."""
        priv = _u8(priv); n = priv.shape[0]
        msgs, off, fixed = self._msgs(msgs, off, n)
        sig = np.zeros((n, 64), np.uint8)
        self.last_seconds = self.lib.drv_ed25519_sign(_p(sig), _p(priv), _p(msgs), None if off is None else _p(off),
                                                      fixed, n, threads)
        return sig

    def ed25519_verify(self, sig, pk, msgs, off=None, threads=1):
        """-> ok[n] int32 (1 valid / 0 invalid); ed25519_VerifySignature, ed25519_verify.c:163."""
        sig = _u8(sig); pk = _u8(pk); n = sig.shape[0]
        msgs, off, fixed = self._msgs(msgs, off, n)
        ok = np.zeros(n, np.int32)
        self.last_seconds = self.lib.drv_ed25519_verify(_p(ok), _p(sig), _p(pk), _p(msgs),
                                                        None if off is None else _p(off), fixed, n, threads)
        return ok
