"""ctypes binding of libcurve25519_b200.so (include/c25519_b200.h).  Fails loudly when the library is
missing or the device is unusable -- there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcurve25519_b200.so")

_lib = None


class EngineError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError("%s not found: build it with `python -m curve25519_b200.build` "
                          "(the engine has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    sigs = {
        "c25519_init": ([i32], i32),
        "c25519_shutdown": ([], i32),
        "c25519_last_error": ([], C.c_char_p),
        "c25519_launch_count": ([], u64),
        "c25519_x25519_shared_batch": ([vp, vp, vp, sz, vp], i32),
        "c25519_x25519_public_batch": ([vp, vp, sz, i32, vp], i32),
        "c25519_x25519_shared_batch_scatter": ([vp, i32, i32, vp, vp, sz, vp], i32),
        "c25519_x25519_shared_kdf_batch": ([vp, sz, vp, vp, sz, vp], i32),
        "c25519_x25519_scalarmult_raw_batch": ([vp, vp, vp, sz, vp], i32),
        "c25519_x25519_scalarmult_raw_host": ([vp, vp, vp, sz], i32),
        "c25519_x25519_shared_host": ([vp, vp, vp, sz], i32),
        "c25519_x25519_public_host": ([vp, vp, sz, i32], i32),
        "c25519_ed25519_keypair_batch": ([vp, vp, vp, sz, vp], i32),
        "c25519_ed25519_sign_batch": ([vp, vp, vp, vp, sz, sz, vp], i32),
        "c25519_ed25519_verify_batch": ([vp, vp, vp, vp, vp, sz, sz, vp], i32),
        "c25519_ed25519_keypair_host": ([vp, vp, vp, sz], i32),
        "c25519_ed25519_sign_host": ([vp, vp, vp, vp, sz, sz], i32),
        "c25519_ed25519_verify_host": ([vp, vp, vp, vp, vp, sz, sz], i32),
        "c25519_ed25519_verify_init_batch": ([vp, vp, sz, vp], i32),
        "c25519_ed25519_verify_check_batch": ([vp, vp, vp, vp, vp, vp, sz, sz, vp], i32),
        "c25519_x25519_shared_kdf_host": ([vp, sz, vp, vp, sz], i32),
        "c25519_modl_batch": ([i32, vp, vp, vp, sz, vp], i32),
        "c25519_modl_host": ([i32, vp, vp, vp, sz], i32),
        "c25519_x25519_shared_sharded": ([vp, vp, vp, sz, vp, vp], i32),
        "c25519_x25519_public_sharded": ([vp, vp, sz, i32, vp, vp], i32),
        "c25519_ed25519_sign_sharded": ([vp, vp, vp, vp, sz, sz, vp, vp], i32),
        "c25519_ed25519_verify_sharded": ([vp, vp, vp, vp, vp, sz, sz, vp, vp], i32),
        "c25519_allgather_records": ([vp, sz, sz, vp, vp], i32),
        "c25519_sharded_register": ([vp, sz, vp], i32),
        "c25519_sharded_unregister": ([vp], i32),
        "c25519_sharded_set_deferred": ([vp, i32], i32),
        "c25519_sharded_sync": ([vp, vp], i32),
        "c25519_nccl_unique_id": ([vp], i32),
        "c25519_nccl_comm_init": ([C.POINTER(vp), i32, i32, vp, i32], i32),
        "c25519_nccl_comm_destroy": ([vp], i32),
        "c25519_test_primitive": ([i32, vp, vp, vp, sz, vp], i32),
        "c25519_imad_peak_kernel": ([C.POINTER(u64), vp, i32, vp], i32),
        # the reference's 11-function API (include/c25519_legacy.h)
        "ecp_TrimSecretKey": ([vp], None),
        "ecp_PointMultiply": ([vp, vp, vp, i32], None),
        "curve25519_dh_CalculatePublicKey": ([vp, vp], None),
        "curve25519_dh_CalculatePublicKey_fast": ([vp, vp], None),
        "curve25519_dh_CreateSharedKey": ([vp, vp, vp], None),
        "ed25519_CreateKeyPair": ([vp, vp, vp, vp], None),
        "ed25519_SignMessage": ([vp, vp, vp, vp, sz], None),
        "ed25519_Blinding_Init": ([vp, vp, sz], vp),
        "ed25519_Blinding_Finish": ([vp], None),
        "ed25519_VerifySignature": ([vp, vp, vp, sz], i32),
        "ed25519_Verify_Init": ([vp, vp], vp),
        "ed25519_Verify_Check": ([vp, vp, vp, sz], i32),
        "ed25519_Verify_Finish": ([vp], None),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(L, name)          # AttributeError here == the .so does not export what the header declares
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


EXPORTED_SYMBOLS = None


def check(rc, what):
    if rc != 0:
        raise EngineError("%s failed (%d): %s" % (what, rc, lib().c25519_last_error().decode()))
