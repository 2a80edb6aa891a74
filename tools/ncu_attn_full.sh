#!/bin/bash
# full ncu capture (with source-level stall sampling) of the attention kernel at the ViT-B/14@518 shape
LIB=${1:-ucod_dpl_b200/csrc/libucod_b200.so}
OUT=${2:-gpurun_out/r2_attn_full}
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 3 -c 1 -f -o $OUT python tools/att_timeline_run.py $LIB > gpurun_out/ncu_attn_full.log 2>&1
tail -n 3 gpurun_out/ncu_attn_full.log
