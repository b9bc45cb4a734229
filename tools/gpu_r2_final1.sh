#!/bin/bash
# round-2 final single-GPU evidence: bench (both arms), launch list, ncu --set full of the three big kernels, sanitizer
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --metric ed25519_verify --steps 10 --warmup 3 --no-secondary > gpurun_out/r2_bench_n1_verify_metric.json 2> gpurun_out/r2_bench_n1_verify_metric.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_x25519_ladder -s 2 -c 1 -f -o gpurun_out/r2_ladder python tools/prof_ladder.py 1048576 shared > gpurun_out/r2_ncu_ladder.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ed25519_verify' -s 2 -c 2 -f -o gpurun_out/r2_verify python tools/prof_ladder.py 1048576 verify > gpurun_out/r2_ncu_verify.log 2>&1
bash tools/sanitize.sh > gpurun_out/r2_compute_sanitizer.txt 2>&1
echo done
