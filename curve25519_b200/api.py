"""Host-side mirror of the reference's operator interface, batched.

Every function accepts either
  * numpy uint8 arrays (HOST buffers): routed through the c25519_*_host entry points, i.e. the call
    includes H2D staging, the kernels and D2H staging -- this is the reference-facing path `bench.py`
    times as `e2e`; or
  * torch uint8 CUDA tensors (DEVICE buffers, HBM-resident): routed through the c25519_*_batch entry
    points on torch's current stream, asynchronous, no copies.
Argument meaning and error behaviour follow the reference's C API (file:line cited per function):
buffers are fixed-size records, secret keys are clamped in place (the clamped copy is returned), there is
no input validation and no error channel besides EngineError for CUDA failures.
"""
import ctypes as C

import numpy as np

from ._native import check, lib

try:  # torch is plumbing (device memory + streams); the numpy/host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_dev(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _np(a, rec=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if rec is not None and (a.ndim != 2 or a.shape[1] != rec):
        raise ValueError("expected uint8 array of shape [n, %d], got %s" % (rec, a.shape))
    return a


def _tt(t, rec=None):
    if t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous torch.uint8 CUDA tensor")
    if rec is not None and (t.dim() != 2 or t.shape[1] != rec):
        raise ValueError("expected shape [n, %d], got %s" % (rec, tuple(t.shape)))
    return t


def _np_out(a, n, rec):
    """A caller-supplied HOST output array is handed to native code as is: it must already be exactly what the C ABI
    writes -- C-contiguous uint8 [n, rec] (no silent copy: the caller would never see the result)."""
    if not isinstance(a, np.ndarray) or a.dtype != np.uint8 or a.shape != (n, rec) or not a.flags["C_CONTIGUOUS"] \
            or not a.flags["WRITEABLE"]:
        raise ValueError("out must be a writable C-contiguous uint8 array of shape [%d, %d]" % (n, rec))
    return a


def _p(a):
    if _is_dev(a):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def init(device=0):
    check(lib().c25519_init(int(device)), "c25519_init")


def launch_count():
    return int(lib().c25519_launch_count())


# ---------------------------------------------------------------------------------------------- X25519
def x25519_shared(pk, sk, out=None, sk_inplace=False):
    """curve25519_dh_CreateSharedKey (curve25519_dh.c:201) over a batch.

    Returns (shared[n,32], clamped_sk[n,32]).  The reference clamps `sk` in place; here a copy is clamped
    unless sk_inplace=True."""
    L = lib()
    if _is_dev(sk):
        pk = _tt(pk, 32); sk = _tt(sk, 32)
        skc = sk if sk_inplace else sk.clone()
        out = torch.empty_like(skc) if out is None else _tt(out, 32)
        if pk.shape[0] != skc.shape[0] or out.shape[0] != skc.shape[0]:
            raise ValueError("pk, sk and out must have the same number of records")
        check(L.c25519_x25519_shared_batch(_p(out), _p(pk), _p(skc), skc.shape[0], _stream()), "x25519_shared_batch")
        return out, skc
    pk = _np(pk, 32); sk = _np(sk, 32)
    skc = sk if sk_inplace else sk.copy()
    out = np.empty_like(skc) if out is None else _np_out(out, skc.shape[0], 32)
    if pk.shape[0] != skc.shape[0]:
        raise ValueError("pk and sk must have the same number of records")
    check(L.c25519_x25519_shared_host(_p(out), _p(pk), _p(skc), skc.shape[0]), "x25519_shared_host")
    return out, skc


def x25519_shared_scatter(peer_buffers, rank, pk, sk, sk_inplace=True):
    """Fused compute + gather: this rank's shared keys are stored by the kernel into row block `rank` of every tensor in
    peer_buffers (the [world*n_local, 32] gathered arrays of all ranks, peer-mapped; e.g. symmetric-memory buffers)."""
    pk = _tt(pk, 32); sk = _tt(sk, 32)
    skc = sk if sk_inplace else sk.clone()
    world = len(peer_buffers)
    ptrs = (C.c_void_p * world)(*[C.c_void_p(b.data_ptr()) for b in peer_buffers])
    check(lib().c25519_x25519_shared_batch_scatter(ptrs, world, int(rank), _p(pk), _p(skc), skc.shape[0], _stream()),
          "x25519_shared_batch_scatter")
    return skc


def x25519_shared_kdf(pk, sk, key_size=32):
    """X25519Private::CreateSharedKey (C++/x25519.cpp:75-95): SHA-512(shared secret)[:key_size] per record."""
    if not _is_dev(sk):
        pk = _np(pk, 32); skc = _np(sk, 32).copy()
        out = np.empty((skc.shape[0], key_size), np.uint8)
        check(lib().c25519_x25519_shared_kdf_host(_p(out), key_size, _p(pk), _p(skc), skc.shape[0]), "x25519_shared_kdf_host")
        return out
    pk = _tt(pk, 32); skc = _tt(sk, 32).clone()
    out = torch.empty((skc.shape[0], key_size), dtype=torch.uint8, device=skc.device)
    check(lib().c25519_x25519_shared_kdf_batch(_p(out), key_size, _p(pk), _p(skc), skc.shape[0], _stream()), "x25519_shared_kdf_batch")
    return out


def x25519_scalarmult_raw(point, scalar):
    """ecp_PointMultiply (curve25519_dh.c:94) over a batch: k*P for ANY 256-bit scalar, no clamping."""
    L = lib()
    if _is_dev(point):
        point = _tt(point, 32); scalar = _tt(scalar, 32)
        out = torch.empty_like(point)
        check(L.c25519_x25519_scalarmult_raw_batch(_p(out), _p(point), _p(scalar), point.shape[0], _stream()), "x25519_scalarmult_raw_batch")
        return out
    point = _np(point, 32); scalar = _np(scalar, 32)
    out = np.empty_like(point)
    check(L.c25519_x25519_scalarmult_raw_host(_p(out), _p(point), _p(scalar), point.shape[0]), "x25519_scalarmult_raw_host")
    return out


def x25519_public(sk, ladder=False, out=None, sk_inplace=False):
    """curve25519_dh_CalculatePublicKey_fast (dh.c:182, ladder=False: 8-fold comb) or
    curve25519_dh_CalculatePublicKey (dh.c:192, ladder=True) over a batch -> (pk[n,32], clamped_sk)."""
    L = lib()
    if _is_dev(sk):
        sk = _tt(sk, 32)
        skc = sk if sk_inplace else sk.clone()
        out = torch.empty_like(skc) if out is None else _tt(out, 32)
        if out.shape[0] != skc.shape[0]:
            raise ValueError("sk and out must have the same number of records")
        check(L.c25519_x25519_public_batch(_p(out), _p(skc), skc.shape[0], int(bool(ladder)), _stream()), "x25519_public_batch")
        return out, skc
    sk = _np(sk, 32)
    skc = sk if sk_inplace else sk.copy()
    out = np.empty_like(skc) if out is None else _np_out(out, skc.shape[0], 32)
    check(L.c25519_x25519_public_host(_p(out), _p(skc), skc.shape[0], int(bool(ladder))), "x25519_public_host")
    return out, skc


# ---------------------------------------------------------------------------------------------- Ed25519
def _msgs_host(msgs, off, n):
    if off is None:
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        if msgs.ndim != 2 or msgs.shape[0] != n:
            raise ValueError("fixed-length messages must be uint8 [n, len]")
        fixed = msgs.shape[1]
        if msgs.size == 0:
            msgs = np.zeros(1, np.uint8)
        return msgs, None, fixed
    off = np.ascontiguousarray(off, dtype=np.uint64)
    if off.shape != (n + 1,):
        raise ValueError("msg_off must have n+1 entries")
    msgs = np.ascontiguousarray(msgs, dtype=np.uint8).reshape(-1)
    if msgs.size == 0:
        msgs = np.zeros(1, np.uint8)
    return msgs, off, 0


def _msgs_dev(msgs, off, n):
    if not _is_dev(msgs) or msgs.dtype != torch.uint8 or not msgs.is_cuda:
        raise ValueError("messages must be a torch.uint8 CUDA tensor")
    if off is None:
        if msgs.dim() != 2 or msgs.shape[0] != n:
            raise ValueError("fixed-length messages must be uint8 [n, len]")
        fixed = msgs.shape[1]
        if msgs.numel() == 0:
            msgs = torch.zeros(1, dtype=torch.uint8, device=msgs.device)
        return msgs.contiguous(), None, fixed
    if not _is_dev(off) or not off.is_cuda or (off.dtype != torch.int64 and off.dtype != torch.uint64) or off.numel() != n + 1:
        raise ValueError("msg_off must be a 64-bit integer CUDA tensor with n+1 entries")
    if msgs.numel() == 0:
        msgs = torch.zeros(1, dtype=torch.uint8, device=off.device)
    return msgs.contiguous(), off.contiguous(), 0


def ed25519_keypair(seed):
    """ed25519_CreateKeyPair (ed25519_sign.c:344) -> (pub[n,32], priv[n,64] = seed || pub)."""
    L = lib()
    if _is_dev(seed):
        seed = _tt(seed, 32); n = seed.shape[0]
        pub = torch.empty((n, 32), dtype=torch.uint8, device=seed.device)
        priv = torch.empty((n, 64), dtype=torch.uint8, device=seed.device)
        check(L.c25519_ed25519_keypair_batch(_p(pub), _p(priv), _p(seed), n, _stream()), "ed25519_keypair_batch")
        return pub, priv
    seed = _np(seed, 32); n = seed.shape[0]
    pub = np.empty((n, 32), np.uint8); priv = np.empty((n, 64), np.uint8)
    check(L.c25519_ed25519_keypair_host(_p(pub), _p(priv), _p(seed), n), "ed25519_keypair_host")
    return pub, priv


def ed25519_sign(priv, msgs, off=None):
    """ed25519_SignMessage (ed25519_sign.c:372) -> sig[n,64].  msgs: [n,len] or flat bytes + off[n+1]."""
    L = lib()
    if _is_dev(priv):
        priv = _tt(priv, 64); n = priv.shape[0]
        msgs, off, fixed = _msgs_dev(msgs, off, n)
        sig = torch.empty((n, 64), dtype=torch.uint8, device=priv.device)
        check(L.c25519_ed25519_sign_batch(_p(sig), _p(priv), _p(msgs), None if off is None else _p(off), fixed, n, _stream()),
              "ed25519_sign_batch")
        return sig
    priv = _np(priv, 64); n = priv.shape[0]
    msgs, off, fixed = _msgs_host(msgs, off, n)
    sig = np.empty((n, 64), np.uint8)
    check(L.c25519_ed25519_sign_host(_p(sig), _p(priv), _p(msgs), None if off is None else _p(off), fixed, n), "ed25519_sign_host")
    return sig


def ed25519_verify(sig, pk, msgs, off=None):
    """ed25519_VerifySignature (ed25519_verify.c:163) -> ok[n] int32, 1 = valid, 0 = invalid."""
    L = lib()
    if _is_dev(sig):
        sig = _tt(sig, 64); pk = _tt(pk, 32); n = sig.shape[0]
        msgs, off, fixed = _msgs_dev(msgs, off, n)
        ok = torch.empty(n, dtype=torch.int32, device=sig.device)
        check(L.c25519_ed25519_verify_batch(_p(ok), _p(sig), _p(pk), _p(msgs), None if off is None else _p(off), fixed, n, _stream()),
              "ed25519_verify_batch")
        return ok
    sig = _np(sig, 64); pk = _np(pk, 32); n = sig.shape[0]
    msgs, off, fixed = _msgs_host(msgs, off, n)
    ok = np.empty(n, np.int32)
    check(L.c25519_ed25519_verify_host(_p(ok), _p(sig), _p(pk), _p(msgs), None if off is None else _p(off), fixed, n),
          "ed25519_verify_host")
    return ok


def ed25519_verify_init(pk):
    """ed25519_Verify_Init (ed25519_verify.c:179) over a batch of keys -> device ctx [n_keys, 2080] uint8."""
    pk = _tt(pk, 32); n = pk.shape[0]
    ctx = torch.empty((n, 2080), dtype=torch.uint8, device=pk.device)
    check(lib().c25519_ed25519_verify_init_batch(_p(ctx), _p(pk), n, _stream()), "ed25519_verify_init_batch")
    return ctx


def ed25519_verify_check(ctx, sig, msgs, off=None, key_index=None):
    """ed25519_Verify_Check (ed25519_verify.c:287): item i is checked against ctx[key_index[i]] (or ctx[i])."""
    sig = _tt(sig, 64); n = sig.shape[0]
    msgs, off, fixed = _msgs_dev(msgs, off, n)
    ok = torch.empty(n, dtype=torch.int32, device=sig.device)
    if key_index is not None and key_index.dtype != torch.int32:
        raise ValueError("key_index must be int32")
    check(lib().c25519_ed25519_verify_check_batch(_p(ok), _p(ctx), None if key_index is None else _p(key_index), _p(sig), _p(msgs),
                                                  None if off is None else _p(off), fixed, n, _stream()),
          "ed25519_verify_check_batch")
    return ok


# ---------------------------------------------------------------------------------------------- mod L
MODL_MULMOD, MODL_ADDMOD, MODL_MONTMUL, MODL_EXPMOD, MODL_INVMOD = range(5)


def modl(op, a, b=None):
    """Arithmetic modulo the group order on 32-byte little-endian records (the self-test's eco_MulMod / eco_AddMod /
    eco_MontMul / eco_ExpModBPO / eco_InvModBPO, test/curve25519_selftest.c:160-282); results canonical."""
    L = lib()
    if _is_dev(a):
        a = _tt(a, 32); b = None if b is None else _tt(b, 32)
        out = torch.empty_like(a)
        check(L.c25519_modl_batch(int(op), _p(out), _p(a), None if b is None else _p(b), a.shape[0], _stream()), "modl_batch")
        return out
    a = _np(a, 32); b = None if b is None else _np(b, 32)
    out = np.empty_like(a)
    check(L.c25519_modl_host(int(op), _p(out), _p(a), None if b is None else _p(b), a.shape[0]), "modl_host")
    return out


# ---------------------------------------------------------------------------------------------- multi-GPU (NCCL inside the C ABI)
class ShardedComm:
    """An ncclComm_t created through the engine's own NCCL binding (c25519_nccl_comm_init), for the *_sharded entry
    points.  `bootstrap(id_bytes_or_None) -> id_bytes` must hand rank 0's 128-byte unique id to every rank (e.g. a
    torch.distributed broadcast); the communicator lives on CUDA device `device`."""

    def __init__(self, world, rank, device, bootstrap):
        L = lib()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            check(L.c25519_nccl_unique_id(buf), "c25519_nccl_unique_id")
        uid = bootstrap(bytes(buf) if rank == 0 else None)
        self.world, self.rank = world, rank
        self.handle = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(L.c25519_nccl_comm_init(C.byref(self.handle), world, rank, idbuf, int(device)), "c25519_nccl_comm_init")

    def register(self, tensor):
        """c25519_sharded_register: results written into `tensor` are exchanged by copy-engine pushes into the peers'
        copies (CUDA IPC) instead of NCCL kernels.  Collective.  Returns False when the memory cannot be shared."""
        rc = lib().c25519_sharded_register(_p(tensor), tensor.numel() * tensor.element_size(), self.handle)
        return rc == 0

    def unregister(self, tensor):
        lib().c25519_sharded_unregister(_p(tensor))

    def set_deferred(self, tensor, on=True):
        """Deferred exchange on a registered array: the transfer to the peers overlaps the caller's next work; sync() joins it."""
        check(lib().c25519_sharded_set_deferred(_p(tensor), 1 if on else 0), "c25519_sharded_set_deferred")

    def sync(self, tensor):
        check(lib().c25519_sharded_sync(_p(tensor), _stream()), "c25519_sharded_sync")

    def close(self):
        if self.handle:
            lib().c25519_nccl_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def x25519_shared_sharded(comm, out_all, pk_local, sk_local):
    """c25519_x25519_shared_sharded: this rank's n_local operations + the NCCL exchange; out_all [world*n_local, 32]."""
    pk_local = _tt(pk_local, 32); sk_local = _tt(sk_local, 32); out_all = _tt(out_all, 32)
    n = sk_local.shape[0]
    if out_all.shape[0] != comm.world * n or pk_local.shape[0] != n:
        raise ValueError("out_all must hold world * n_local records")
    check(lib().c25519_x25519_shared_sharded(_p(out_all), _p(pk_local), _p(sk_local), n, comm.handle, _stream()), "x25519_shared_sharded")
    return out_all


def ed25519_sign_sharded(comm, sig_all, priv_local, msgs_local):
    priv_local = _tt(priv_local, 64); sig_all = _tt(sig_all, 64); n = priv_local.shape[0]
    msgs, off, fixed = _msgs_dev(msgs_local, None, n)
    if sig_all.shape[0] != comm.world * n:
        raise ValueError("sig_all must hold world * n_local records")
    check(lib().c25519_ed25519_sign_sharded(_p(sig_all), _p(priv_local), _p(msgs), None, fixed, n, comm.handle, _stream()), "ed25519_sign_sharded")
    return sig_all


def ed25519_verify_sharded(comm, ok_all, sig_local, pk_local, msgs_local):
    sig_local = _tt(sig_local, 64); pk_local = _tt(pk_local, 32); n = sig_local.shape[0]
    msgs, off, fixed = _msgs_dev(msgs_local, None, n)
    if ok_all.dtype != torch.int32 or ok_all.numel() != comm.world * n or not ok_all.is_contiguous():
        raise ValueError("ok_all must be a contiguous int32 tensor of world * n_local entries")
    check(lib().c25519_ed25519_verify_sharded(_p(ok_all), _p(sig_local), _p(pk_local), _p(msgs), None, fixed, n, comm.handle, _stream()),
          "ed25519_verify_sharded")
    return ok_all


def allgather_records(comm, all_records, n_local):
    """In-place NCCL all-gather of fixed-size records (this rank's block already at row rank * n_local)."""
    rec = all_records.numel() * all_records.element_size() // (comm.world * n_local)
    check(lib().c25519_allgather_records(_p(all_records), rec, n_local, comm.handle, _stream()), "allgather_records")
    return all_records


# ---------------------------------------------------------------------------------------------- test hooks
def test_primitive(op, a, b=None, out_rec=32):
    """Run one device primitive per record (see c25519_test_primitive in include/c25519_b200.h)."""
    a = _tt(a); n = a.shape[0]
    out = torch.empty((n, out_rec), dtype=torch.uint8, device=a.device)
    check(lib().c25519_test_primitive(int(op), _p(out), _p(a), None if b is None else _p(_tt(b)), n, _stream()), "test_primitive")
    return out


def imad_peak(iters=20000, repeats=3, accumulate=False):
    """Measured IMAD.WIDE.U32 rate of this device in MAC32/s, best of `repeats`.
    accumulate=False: fresh form (Rc = RZ), the fastest form and the conservative roofline denominator;
    accumulate=True : 64-bit-accumulator form, the one a multi-precision MAC actually needs."""
    sink = torch.arange(64, dtype=torch.int32, device="cuda") * 0x01010101 + 12345
    macs = C.c_uint64(0)
    L = lib()
    sign = -1 if accumulate else 1
    check(L.c25519_imad_peak_kernel(C.byref(macs), _p(sink), sign * 2000, _stream()), "imad_peak")
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(repeats):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(L.c25519_imad_peak_kernel(C.byref(macs), _p(sink), sign * int(iters), _stream()), "imad_peak")
        e1.record(); torch.cuda.synchronize()
        best = max(best, macs.value / (e0.elapsed_time(e1) * 1e-3))
    return best
