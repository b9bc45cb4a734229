"""Build libcurve25519_b200.so (the CUDA engine + C ABI) in-tree with nvcc for sm_100a.

    python -m curve25519_b200.build [--force] [--verbose]

Each translation unit under csrc/ is compiled to an object (in parallel) with
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and linked into curve25519_b200/libcurve25519_b200.so.  The .so is git-ignored but travels with gpurun
snapshots; nothing is JIT-compiled at import time and nothing falls back to another implementation.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libcurve25519_b200.so")
UNITS = ["engine.cu", "x25519_kernels.cu", "ed25519_kernels.cu", "modl_kernels.cu", "legacy_internals.cu", "test_kernels.cu", "comb_table.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-Wall", "--expt-relaxed-constexpr"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", f) for f in ("c25519_b200.h", "c25519_legacy.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(unit, verbose):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ, unit[:-3] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
    return unit, r.stderr


def build(force=False, verbose=False):
    if not os.path.exists(os.path.join(CSRC, "comb_table.cu")):
        subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "gen_base_table.py")], check=True)
    deps = _deps()
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [d for d in deps if not d.endswith(".cu")]
    todo = [u for u in UNITS if force or _stale(os.path.join(OBJ, u[:-3] + ".o"), [os.path.join(CSRC, u)] + hdrs)]
    with cf.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for unit, log in ex.map(lambda u: _compile(u, verbose), todo):
            if verbose:
                print("==", unit, "\n", log)
    objs = [os.path.join(OBJ, u[:-3] + ".o") for u in UNITS]
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-Bsymbolic", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
