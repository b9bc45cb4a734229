#!/bin/bash
# round-2 GPU call 1: instruction-rate probe, baseline parity, ladder ncu with the north-star metrics
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
nproc >> gpurun_out/r2_c1_smi.txt
./tools/ubench6 > gpurun_out/r2_ubench6.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c1_pytest.txt 2>&1
M=smsp__sass_thread_inst_executed_op_integer_pred_on.sum,smsp__sass_thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,sm__cycles_active.avg,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_elapsed.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_x25519_ladder -s 2 -c 1 --csv --log-file gpurun_out/r2_c1_ladder_metrics.csv python tools/prof_ladder.py 1048576 shared > gpurun_out/r2_c1_ncu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
echo done
