"""GPU: the multi-GPU C entry points (c25519_*_sharded, NCCL bound inside the library).

* world = 1 (any box): the NCCL binding loads, a single-rank communicator is created through the engine's own helpers,
  and the sharded calls reproduce the plain batch calls.
* world = 2 (boxes with >= 2 GPUs): BASELINE config 5 in miniature -- X25519 shared keys + Ed25519 sign + verify, each
  sharded over the ranks, gathered by ONE exchange per call, and the GATHERED buffers are checked on EVERY rank against
  the oracle (no GPU parity check of the gathered buffer existed)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_single_rank_communicator(engine, oracle, rng):
    import torch
    comm = engine.ShardedComm(1, 0, torch.cuda.current_device(), lambda uid: uid)
    try:
        n = (1 << 18) + 77                                          # world == 1: four pipelined slices, no exchange
        sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        d_sk = torch.from_numpy(sk).cuda(); d_pk = torch.from_numpy(pk).cuda()
        out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
        engine.x25519_shared_sharded(comm, out, d_pk, d_sk)
        exp, exp_sk = oracle.x25519_shared(pk[:4096], sk[:4096], threads=os.cpu_count() or 1)
        assert (out[:4096].cpu().numpy() == exp).all() and (d_sk[:4096].cpu().numpy() == exp_sk).all()
        ref, _ = engine.x25519_shared(d_pk, torch.from_numpy(sk).cuda())
        assert torch.equal(out, ref)
        rec = torch.arange(64 * 100, dtype=torch.int64, device="cuda").to(torch.uint8).reshape(100, 64)
        keep = rec.clone()
        engine.allgather_records(comm, rec, 100)
        assert torch.equal(rec, keep)
    finally:
        comm.close()


def _worker(rank, world, port, n_local, q):
    try:
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from curve25519_b200 import api
        from oracle import pyoracle
        o = pyoracle.Oracle("reference" if pyoracle.available("reference") else "port")

        def bootstrap(uid):
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = api.ShardedComm(world, rank, rank, bootstrap)
        n = world * n_local
        rng = np.random.Generator(np.random.PCG64(0xC0F5))           # the same global job on every rank
        sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); pk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        ne = 4096 * world; ne_local = ne // world
        seed = rng.integers(0, 256, (ne, 32), dtype=np.uint8); msgs = rng.integers(0, 256, (ne, 64), dtype=np.uint8)
        T = max(1, (os.cpu_count() or 2) // world)
        lo, hi = rank * n_local, (rank + 1) * n_local
        # --- X25519 shared keys: 4 pipelined slices (inversion + exchange under the next ladder) when n_local >= 2^18
        out_all = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
        d_sk = torch.from_numpy(sk[lo:hi]).cuda()
        api.x25519_shared_sharded(comm, out_all, torch.from_numpy(pk[lo:hi]).cuda(), d_sk)
        torch.cuda.synchronize()
        exp, exp_sk = o.x25519_shared(pk, sk, threads=T)
        assert (out_all.cpu().numpy() == exp).all(), "gathered shared keys differ on rank %d" % rank
        assert (d_sk.cpu().numpy() == exp_sk[lo:hi]).all()
        # the same call on a REGISTERED result array: fused path (the inversion kernel stores into every rank's array through
        # CUDA IPC mappings, stream-memory-op flags, no NCCL kernel); twice, to exercise the epoch protocol
        reg_all = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
        assert comm.register(reg_all), "c25519_sharded_register failed"
        for rep in range(2):
            reg_all.zero_(); torch.cuda.synchronize(); dist.barrier()
            api.x25519_shared_sharded(comm, reg_all, torch.from_numpy(pk[lo:hi]).cuda(), torch.from_numpy(sk[lo:hi]).cuda())
            torch.cuda.synchronize()
            assert (reg_all.cpu().numpy() == exp).all(), "fused gathered shared keys differ on rank %d (rep %d)" % (rank, rep)
        # deferred exchange: three batches back to back (each transfer under the next ladder), joined once at the end
        comm.set_deferred(reg_all, True)
        rng2 = np.random.Generator(np.random.PCG64(0xDEF))
        batches = [(rng2.integers(0, 256, (n, 32), dtype=np.uint8), rng2.integers(0, 256, (n, 32), dtype=np.uint8)) for _ in range(3)]
        for bsk, bpk in batches:
            api.x25519_shared_sharded(comm, reg_all, torch.from_numpy(bpk[lo:hi]).cuda(), torch.from_numpy(bsk[lo:hi]).cuda())
        comm.sync(reg_all); torch.cuda.synchronize(); dist.barrier()
        exp3, _ = o.x25519_shared(batches[2][1], batches[2][0], threads=T)
        assert (reg_all.cpu().numpy() == exp3).all(), "deferred gathered shared keys differ on rank %d" % rank
        comm.set_deferred(reg_all, False)
        # generic records on the registered array: copy-engine pushes
        reg_all.zero_(); reg_all[lo:hi] = torch.from_numpy(exp[lo:hi]).cuda(); torch.cuda.synchronize(); dist.barrier()
        api.allgather_records(comm, reg_all, n_local)
        torch.cuda.synchronize()
        assert (reg_all.cpu().numpy() == exp).all(), "pushed records differ on rank %d" % rank
        dist.barrier()
        comm.unregister(reg_all)
        # --- Ed25519 sign + verify, sharded
        e_pub, e_priv = o.ed25519_keypair(seed, threads=T)
        elo, ehi = rank * ne_local, (rank + 1) * ne_local
        sig_all = torch.zeros((ne, 64), dtype=torch.uint8, device="cuda")
        api.ed25519_sign_sharded(comm, sig_all, torch.from_numpy(e_priv[elo:ehi]).cuda(), torch.from_numpy(msgs[elo:ehi]).cuda())
        torch.cuda.synchronize()
        e_sig = o.ed25519_sign(e_priv, msgs, threads=T)
        assert (sig_all.cpu().numpy() == e_sig).all(), "gathered signatures differ on rank %d" % rank
        bad = e_sig.copy(); bad[::5, 3] ^= 0x10
        ok_all = torch.full((ne,), -1, dtype=torch.int32, device="cuda")
        api.ed25519_verify_sharded(comm, ok_all, torch.from_numpy(bad[elo:ehi]).cuda(), torch.from_numpy(e_pub[elo:ehi]).cuda(),
                                   torch.from_numpy(msgs[elo:ehi]).cuda())
        torch.cuda.synchronize()
        assert (ok_all.cpu().numpy() == o.ed25519_verify(bad, e_pub, msgs, threads=T)).all(), "gathered verdicts differ"
        # --- config 5's packing: uniform 64-byte records gathered once
        rec = torch.zeros((world * 3 * ne_local, 64), dtype=torch.uint8, device="cuda")
        mine = rec[rank * 3 * ne_local:(rank + 1) * 3 * ne_local]
        mine[:ne_local, :32] = out_all[lo:lo + ne_local]
        mine[ne_local:2 * ne_local] = sig_all[elo:ehi]
        mine[2 * ne_local:, :4] = ok_all[elo:ehi].view(torch.uint8).view(ne_local, 4)
        api.allgather_records(comm, rec, 3 * ne_local)
        torch.cuda.synchronize()
        h = rec.cpu().numpy()
        for r in range(world):
            blk = h[r * 3 * ne_local:(r + 1) * 3 * ne_local]
            assert (blk[:ne_local, :32] == exp[r * n_local:r * n_local + ne_local]).all()
            assert (blk[ne_local:2 * ne_local] == e_sig[r * ne_local:(r + 1) * ne_local]).all()
        comm.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as ex:                                          # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + repr(ex) + "\n" + traceback.format_exc()))


def test_config5_gathered_buffers_on_every_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    ps = [ctx.Process(target=_worker, args=(r, world, port, (1 << 18) + 512, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=600) for _ in ps]
    [p.join(60) for p in ps]
    assert all(r[1] == "ok" for r in res), res
