"""CPU-only: disassemble the built objects and pin the instruction mix the roofline numbers rest on.

Round 1's "fresh IMAD.WIDE peak" loop executed ONE multiply per trip (ptxas hoisted the other seven), so its
number was not a multiply rate.  These tests make that impossible to regress silently: they
run `cuobjdump -sass` on curve25519_b200/csrc/_obj/*.o (nvcc and cuobjdump exist in the authoring image and on the
GPU boxes) and count mnemonics inside each kernel's hot loop with tools/sass_mix.py.
"""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not os.path.exists("/usr/local/cuda/bin/cuobjdump") and not shutil.which("cuobjdump"),
                                reason="cuobjdump not available")


@pytest.fixture(scope="module")
def objs():
    from curve25519_b200 import build
    build.build()
    import sass_mix
    d = os.path.join(ROOT, "curve25519_b200", "csrc", "_obj")
    return sass_mix, {n: sass_mix.kernels(os.path.join(d, n + ".o")) for n in ("test_kernels", "x25519_kernels", "ed25519_kernels")}


def _loop_mix(sass_mix, ins):
    hl = sass_mix.hot_loop(ins)
    assert hl is not None, "kernel has no loop"
    return sass_mix.mix(ins, *hl)


def _find(kernels, needle):
    hits = [k for k in kernels if needle in k]
    assert hits, "kernel %s not found in %s" % (needle, list(kernels))
    return hits


def test_peak_kernels_execute_eight_wide_multiplies_per_trip(objs):
    sass_mix, o = objs
    ks = o["test_kernels"]
    fresh = ks[_find(ks, "k_imad_peakILi0E")[0]]
    acc = ks[_find(ks, "k_imad_peakILi1E")[0]]
    for name, ins in (("fresh", fresh), ("accumulate", acc)):
        m = _loop_mix(sass_mix, ins)
        wide = sum(n for mn, n in m.items() if mn.startswith("IMAD.WIDE"))
        assert wide == 8, "%s peak loop holds %d IMAD.WIDE per trip, credited with 8: %s" % (name, wide, dict(m))
        other_fma = sum(n for mn, n in m.items() if mn.startswith("IMAD") and not mn.startswith("IMAD.WIDE"))
        assert other_fma == 0, "%s peak loop has other multiply-pipe instructions: %s" % (name, dict(m))
    # operand forms: fresh = zero accumulator (RZ), accumulate = Rd is also the addend
    hl = sass_mix.hot_loop(fresh)
    for addr, mn, ops in fresh:
        if hl[0] <= addr <= hl[1] and mn.startswith("IMAD.WIDE"):
            assert ops.rstrip().endswith("RZ"), ops
    hl = sass_mix.hot_loop(acc)
    for addr, mn, ops in acc:
        if hl[0] <= addr <= hl[1] and mn.startswith("IMAD.WIDE"):
            r = [x.strip() for x in ops.split(",")]
            assert r[0] == r[-1] and r[-1] != "RZ", ops


def test_ladder_loop_multiply_count_matches_the_algorithm(objs):
    """One ladder step = 5 M + 4 S + 1 W: 5*73 + 4*45 + 9 = 554 multiply instructions (IMAD.WIDE + IMAD + IMAD.HI);
    the shipped loop must not execute materially more 64-bit multiplies than that."""
    sass_mix, o = objs
    ks = o["x25519_kernels"]
    name = _find(ks, "k_x25519_ladderILb1E")[0]
    m = _loop_mix(sass_mix, ks[name])
    wide = sum(n for mn, n in m.items() if mn.startswith("IMAD.WIDE"))
    assert 500 <= wide <= 560, dict(m)
    assert not any(mn.startswith(("LDL", "STL")) for mn in m), "ladder loop spills: %s" % dict(m)


def test_tma_bulk_copy_stages_the_comb_table(objs):
    sass_mix, o = objs
    ks = o["ed25519_kernels"]
    for needle in ("k_x25519_comb", "k_ed25519_keypair", "k_ed25519_verify"):
        hit = [k for k in _find(ks, needle) if any(mn.startswith("UBLKCP") for _, mn, _ in ks[k])]
        assert hit, "no UBLKCP (cp.async.bulk) in any %s kernel" % needle
