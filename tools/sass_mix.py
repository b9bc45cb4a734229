#!/usr/bin/env python3
"""sass_mix.py -- count SASS mnemonics inside the loop(s) of each kernel of a cubin / object / executable.

    python tools/sass_mix.py <file> [kernel-name-regex]

For every kernel whose (mangled) name matches the regex, the instruction stream printed by `cuobjdump -sass`
is scanned for backward branches; each backward branch closes a loop [target, branch].  The LARGEST loop that
contains no other backward-branch target is reported per kernel ("hot loop") together with the whole-kernel
mix.  Used by tests/test_sass.py (no GPU needed) to pin that the roofline micro-kernels really execute the
instructions they are credited with, and by hand to document the instruction mix of the shipped kernels.
"""
import collections
import re
import subprocess
import sys

CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")


def kernels(path):
    """-> {kernel_name: [(addr, mnemonic, operands), ...]}"""
    out = subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            res[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
    return res


def loops(instrs):
    """-> list of (start_addr, end_addr) for every backward branch."""
    res = []
    for addr, mn, ops in instrs:
        if mn.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", ops)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= addr:
                    res.append((tgt, addr))
    return res


def mix(instrs, lo=None, hi=None):
    c = collections.Counter()
    for addr, mn, _ in instrs:
        if lo is not None and not (lo <= addr <= hi):
            continue
        c[mn] += 1
    return c


def hot_loop(instrs):
    ls = loops(instrs)
    if not ls:
        return None
    inner = [l for l in ls if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in ls)]
    return max(inner, key=lambda l: l[1] - l[0])


def group(c):
    """Collapse modifiers: IMAD.WIDE.U32.X -> IMAD.WIDE ; IADD3.X -> IADD3 ; keep IMAD.X / IMAD.MOV / IMAD.IADD / IMAD.SHL apart."""
    g = collections.Counter()
    for mn, n in c.items():
        parts = mn.split(".")
        key = parts[0]
        if key == "IMAD" and len(parts) > 1 and parts[1] in ("WIDE", "HI", "X", "MOV", "IADD", "SHL", "U32"):
            key = "IMAD" if parts[1] == "U32" else "IMAD." + parts[1]
        g[key] += n
    return g


def main():
    path = sys.argv[1]
    rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    for name, ins in kernels(path).items():
        if rx and not rx.search(name):
            continue
        hl = hot_loop(ins)
        whole = group(mix(ins))
        print("%s: %d instructions" % (name, len(ins)))
        if hl:
            g = group(mix(ins, *hl))
            n = sum(g.values())
            print("  hot loop [%#x, %#x]: %d instr: %s" % (hl[0], hl[1], n, ", ".join("%s=%d" % kv for kv in g.most_common(14))))
        print("  whole: %s" % ", ".join("%s=%d" % kv for kv in whole.most_common(14)))


if __name__ == "__main__":
    main()
