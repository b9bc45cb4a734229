"""GPU: the reference's OWN test program (test/curve25519_test.c: dh_test, signature_test with the RFC 8032
TEST 2 vector with and without blinding, the donna cross-check and speed_test) compiled unmodified in the
authoring container and linked against libcurve25519_b200.so instead of libcurve25519.a.  Its main() is
replaced by oracle/dropin_main.c, which runs dh_test(), signature_test() and the donna cross-check from the top
of speed_test() but not speed_test()'s rdtsc timing loops; exit status is the reference's own failure count."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "curve25519_test_b200")


def test_reference_test_program_passes_against_our_library(engine):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/curve25519_test_b200 not built (needs /root/reference at build time)")
    os.chmod(EXE, 0o755)
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert "FAILED" not in r.stdout.upper() or "0 FAILED" in r.stdout.upper(), tail


def _run(name, timeout=300):
    exe = os.path.join(os.path.dirname(EXE), name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s not built (needs /root/reference at build time)" % name)
    os.chmod(exe, 0o755)
    return subprocess.run([exe], capture_output=True, text=True, timeout=timeout)


def test_reference_selftest_suite_runs_on_the_engine(engine):
    """test/curve25519_selftest.c built with -DECP_SELF_TEST: curve25519_SelfTest(0) and ed25519_selftest()
    drive the library's internal ecp_/eco_/edp_/SHA512_ symbols, every arithmetic one executed on the GPU
    (csrc/legacy_internals.cu).  The same program linked against the compiled reference gives the same verdict."""
    r = _run("curve25519_selftest_b200")
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "curve25519_SelfTest(0): 0 failure(s)" in r.stdout and "ed25519_selftest(): 0 failure(s)" in r.stdout
    assert "FAILED" not in r.stdout
    ref = _run("curve25519_selftest_ref")
    assert ref.returncode == 0 and ref.stdout == r.stdout


def test_reference_cxx_wrappers_against_our_library(engine):
    """C++/x25519.cpp and C++/ed25519.cpp compiled where they lie: X25519Private::CreateShare /
    CreateSharedKey (SHA-512 KDF), ED25519Private::SignMessage (with the build-time blinding contexts),
    ED25519Public::VeifySignature -- byte-identical output whether linked against the reference or against the engine."""
    ours = _run("cxx_dropin_b200")
    assert ours.returncode == 0, (ours.stdout + ours.stderr)[-2000:]
    ref = _run("cxx_dropin_ref")
    assert ref.returncode == 0
    assert ours.stdout == ref.stdout
    assert "verify=1 tampered=0" in ours.stdout
