#!/bin/bash
# compute-sanitizer passes over the small parity tests (run under gpurun).  memcheck + racecheck (the TMA-staged
# shared-memory table and its mbarrier) + initcheck on the scratch records.
set -u
mkdir -p gpurun_out
SEL='kat or low_order or golden or rfc8032 or empty or legacy or generic or kdf or abi or long or two_phase or sc_muladd'
for tool in memcheck racecheck initcheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_x25519.py tests/test_gpu_ed25519.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_$tool.log | tail -5
done
