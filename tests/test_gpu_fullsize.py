"""GPU: 100 % byte equality with the REFERENCE at BASELINE's full sizes (2^20 records, configs 2, 3, 4) on any box.

The expected outputs were computed once by the compiled reference (oracle/_ref) over the seeded inputs of
tests/fullsize.py and are committed as SHA-256 digests (tests/golden/fullsize_digests.json, one digest per block of
4096 records + one per array); no host cores are needed here, so nothing is sampled."""
import json
import os

import numpy as np
import pytest

from . import fullsize as F

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def expected():
    return json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_config2_shared_keys_full_size(engine, expected):
    sk, pk = F.config2_inputs()
    out, skc = engine.x25519_shared(_dev(pk), _dev(sk))
    F.assert_digest("config 2 shared keys", out.cpu().numpy(), expected["config2_shared"])
    F.assert_digest("config 2 clamped secret keys", skc.cpu().numpy(), expected["config2_sk_clamped"])
    # and through the host-pointer ABI (what the legacy wrappers call)
    out_h, skc_h = engine.x25519_shared(pk[: 1 << 18], sk[: 1 << 18])
    assert F.digests(out_h)["blocks"] == expected["config2_shared"]["blocks"][: (1 << 18) // F.BLOCK]


def test_config3_public_keys_full_size(engine, expected):
    sk, _ = F.config2_inputs()
    pub, _ = engine.x25519_public(_dev(sk), ladder=False)
    F.assert_digest("config 3 public keys (comb)", pub.cpu().numpy(), expected["config3_public"])
    pub_l, _ = engine.x25519_public(_dev(sk[: 1 << 17]), ladder=True)
    assert F.digests(pub_l.cpu().numpy())["blocks"] == expected["config3_public"]["blocks"][: (1 << 17) // F.BLOCK]


def test_config4_keygen_sign_verify_full_size(engine, expected):
    seed, msgs = F.config4_inputs()
    pub, priv = engine.ed25519_keypair(_dev(seed))
    F.assert_digest("config 4 public keys", pub.cpu().numpy(), expected["config4_pub"])
    F.assert_digest("config 4 private keys", priv.cpu().numpy(), expected["config4_priv"])
    sig = engine.ed25519_sign(priv, _dev(msgs))
    sig_h = sig.cpu().numpy()
    F.assert_digest("config 4 signatures", sig_h, expected["config4_sig"])
    tsig, tmsgs = F.config4_tamper(sig_h, msgs)
    ok = engine.ed25519_verify(_dev(tsig), pub, _dev(tmsgs)).cpu().numpy()
    assert int(ok.sum()) == expected["config4_ok_count"]
    F.assert_digest("config 4 verdicts", ok, expected["config4_ok"])
    # both outcomes occur, and the S + L items are among the accepted ones
    assert 0 < int(ok.sum()) < F.N_FULL
    assert ok[8] == 1 and ok[0] == 0
