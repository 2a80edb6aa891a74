#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench_first.log | tail -5
