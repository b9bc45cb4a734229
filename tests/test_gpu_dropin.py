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
