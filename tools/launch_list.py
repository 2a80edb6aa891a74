"""Print one step's kernels from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_list.py <csv> [first-kernel-substring] [last-kernel-substring]"""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ni, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
seq = [(int(r[ii]), r[ni].split("(")[0][-52:], float(r[vi].replace(",", "")) / 1e3) for r in rows[2:]]
first = sys.argv[2] if len(sys.argv) > 2 else "upsample_bilinear"
last = sys.argv[3] if len(sys.argv) > 3 else "adamw"
start = next(i for i, s in enumerate(seq) if first in s[1])
tot = 0.0
for s in seq[start:]:
    print(f"{s[0]:4d} {s[1]:54s} {s[2]:8.1f} us")
    tot += s[2]
    if last in s[1]:
        break
print(f"sum {tot:.1f} us")
