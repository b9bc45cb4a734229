#!/bin/bash
# round-2 GPU call 3: new bench line (both headlines), ncu metrics for the verify kernels, launch list
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c3_bench.json 2> gpurun_out/r2_c3_bench.err
M=smsp__sass_thread_inst_executed_op_integer_pred_on.sum,smsp__sass_thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__waves_per_multiprocessor,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_ed25519_verify|k_normalize' -s 3 -c 3 --csv --log-file gpurun_out/r2_c3_verify_metrics.csv python tools/prof_ladder.py 1048576 verify > gpurun_out/r2_c3_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c3_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/r2_c3_bench_under_ncu.log 2>&1
echo done
