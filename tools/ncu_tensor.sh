#!/bin/bash
# tensor-pipe / MUFU utilisation of the attention and GEMM kernels (ncu counters; one GPU)
M=sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.max,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,gpu__time_duration.sum,sm__cycles_active.avg
OUT=${1:-gpurun_out/r2_ncu_tensor.csv}
ncu --metrics $M --clock-control none -k regex:"attention_fwd|gemm2_bf16" -s 120 -c 12 --csv --log-file $OUT python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tensor_stdout.log 2>&1
tail -n 3 gpurun_out/ncu_tensor_stdout.log | cut -c1-300
