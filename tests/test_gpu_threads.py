"""GPU: the C ABI is reentrant like the reference (SURVEY 8b "Threading") and leaves the caller's CUDA device alone."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_host_callers(engine, oracle):
    """Eight threads call the host-pointer ABI at once (more than the per-device pipeline pool holds): every result is
    bit-exact and no caller sees another's data."""
    rngs = [np.random.Generator(np.random.PCG64(100 + t)) for t in range(8)]
    work = []
    for r in rngs:
        n = int(r.integers(300, 3000))
        work.append((r.integers(0, 256, (n, 32), dtype=np.uint8), r.integers(0, 256, (n, 32), dtype=np.uint8),
                     r.integers(0, 256, (n, 48), dtype=np.uint8)))
    res = [None] * 8

    def run(t):
        sk, pk, msgs = work[t]
        out, skc = engine.x25519_shared(pk, sk)
        pub, priv = engine.ed25519_keypair(sk)
        sig = engine.ed25519_sign(priv, msgs)
        ok = engine.ed25519_verify(sig, pub, msgs)
        res[t] = (out, skc, pub, sig, ok)
    th = [threading.Thread(target=run, args=(t,)) for t in range(8)]
    [x.start() for x in th]; [x.join() for x in th]
    for t in range(8):
        sk, pk, msgs = work[t]
        out, skc, pub, sig, ok = res[t]
        e_out, e_sk = oracle.x25519_shared(pk, sk, threads=4)
        assert (out == e_out).all() and (skc == e_sk).all()
        e_pub, e_priv = oracle.ed25519_keypair(sk, threads=4)
        assert (pub == e_pub).all() and (sig == oracle.ed25519_sign(e_priv, msgs, threads=4)).all() and ok.all()


def test_current_device_is_restored(engine):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(1)
    try:
        z = np.zeros((4, 32), np.uint8)
        engine.x25519_shared(z, z)                                  # host path runs on the engine's default device (0)
        assert torch.cuda.current_device() == 1
        a = torch.zeros((4, 32), dtype=torch.uint8, device="cuda:0")
        engine.x25519_shared(a, a.clone())                          # device path follows the pointers
        assert torch.cuda.current_device() == 1
        b = torch.zeros((4, 32), dtype=torch.uint8, device="cuda:1")
        with pytest.raises(Exception, match="different devices"):
            engine.x25519_shared(a, b)
        out, _ = engine.x25519_shared(b, b.clone())                 # second device initialised lazily, same results
        assert out.device.index == 1
    finally:
        torch.cuda.set_device(0)


def test_caller_supplied_host_out_is_validated(engine):
    z = np.zeros((4, 32), np.uint8)
    with pytest.raises(ValueError):
        engine.x25519_shared(z, z, out=np.zeros((4, 16), np.uint8))
    with pytest.raises(ValueError):
        engine.x25519_shared(z, z, out=np.zeros((8, 32), np.uint8)[::2])
    with pytest.raises(ValueError):
        engine.x25519_public(z, out=np.zeros((4, 32), np.int32))
