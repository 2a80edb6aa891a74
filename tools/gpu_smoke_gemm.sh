#!/bin/bash
# first GPU bring-up: GEMM parity + a quick throughput probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/gemm_test.log
timeout 120 python tools/probe_gemm.py 2>&1 | tee gpurun_out/gemm_probe.log
