#!/bin/bash
# `ncu --set full` evidence for the bench workloads, three bounded passes (one GPU).  The reports stay in /tmp on the
# GPU box (they exceed what gpurun brings back); only the summaries land in gpurun_out/ and are copied to profiles/.
#   pass A: two steady-state ViT layers of the Look-Twice step (GEMMs, attention, LayerNorm)
#   pass B: everything else in one Look-Twice step (patch embed, decoder, boxes, crop, paste, upsample)
#   pass C: the side workloads (pseudo-label scorer, training step incl. the tcgen05 weight gradient, CORAL selection)
set -u
OUT=${1:-gpurun_out}
COMMON="--set full --clock-control none -f"
ncu $COMMON -k regex:"gemm2_bf16|gemm_bf16|attention_fwd|layernorm_bf16" -s 60 -c 16 -o /tmp/ncu_a \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-workloads > $OUT/ncu_a.log 2>&1
ncu $COMMON -k regex:"im2col|cls_|decoder_|ccl_|lt_|crop_|paste_|resample_|fill_|upsample|mask_scale|to_tensor|cast_bf16" -c 40 -o /tmp/ncu_b \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-workloads > $OUT/ncu_b.log 2>&1
ncu $COMMON -k regex:"pseudo_label|refine_|wgrad|entropy|range_flag|window_|ge_|apm_|disc|adamw|decoder_bwd|decoder_gram|decoder_ortho|train_loss" -c 60 -o /tmp/ncu_c \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_c.log 2>&1
ls -la /tmp/ncu_*.ncu-rep
python tools/ncu_traffic.py $OUT /tmp/ncu_a.ncu-rep /tmp/ncu_b.ncu-rep /tmp/ncu_c.ncu-rep > $OUT/ncu_traffic.log 2>&1
tail -n 60 $OUT/ncu_traffic.log | cut -c1-230
