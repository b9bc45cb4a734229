#!/bin/bash
# compute-sanitizer passes over the small parity tests (run under gpurun).  memcheck + racecheck (the TMA-staged
# shared-memory table and its mbarrier, the 4-lane cooperative ladder) + initcheck on the scratch records.
set -u
mkdir -p gpurun_out
SEL='kat or low_order or golden or rfc8032 or empty or legacy or generic or kdf or abi or long or two_phase or sc_muladd or modl or selftest_mod or split_key or wide_scalars or ragged'
for tool in memcheck racecheck initcheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_x25519.py tests/test_gpu_ed25519.py tests/test_gpu_modl.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_$tool.log | tail -5
done
echo "=== compute-sanitizer --tool memcheck: the reference's self-test and C++ wrappers on the engine (legacy internals)"
for exe in curve25519_selftest_b200 cxx_dropin_b200 curve25519_test_b200; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 oracle/_ref/$exe > gpurun_out/sanitize_$exe.log 2>&1
  echo "$exe rc=$?"; grep -E "ERROR SUMMARY|all checks passed|failures = " gpurun_out/sanitize_$exe.log | tail -3
done
