#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of bench.py into the JSON bench.py reads for `roofline.traffic`
(profiles/r02_ncu_traffic.json) and a readable per-kernel table (profiles/r02_ncu_kernels.txt).
usage: python tools/ncu_traffic.py <outdir> a.ncu-rep [b.ncu-rep ...]   (see tools/ncu_capture.sh)"""
import csv
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CLASS = [("gemm", ("gemm_bf16_tcgen05", "gemm2_bf16_tcgen05")), ("attention", ("attention_fwd", "attention_rows")),
         ("layernorm", ("layernorm",)), ("embed", ("im2col", "cls_init", "cls_row")),
         ("decoder", ("decoder_",)), ("ccl", ("ccl_", "lt_boxes", "lt_build", "lt_init")),
         ("resample", ("upsample", "crop_", "paste_", "resample_", "fill_", "mask_scale", "to_tensor")),
         ("pseudo_label", ("pseudo_", "pl_", "refine_")), ("coral", ("entropy", "coral_", "window_")),
         ("train", ("apm_", "discriminator", "disc_", "adamw", "ema_", "bce_", "wgrad_"))]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "sm__cycles_elapsed.avg",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def main():
    outdir, reps = Path(sys.argv[1]), sys.argv[2:]
    per = defaultdict(list)
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
        if len(rows) < 3:
            print(f"{rep}: no kernels captured")
            continue
        collect(rows, per)
    report(per, outdir, reps)


def collect(rows, per):
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    name_i = ix["Kernel Name"]
    for r in rows[2:]:
        nm = r[name_i]
        cls = next((c for c, pats in CLASS if any(p in nm for p in pats)), "other")
        d = {}
        for w in WANT:
            j = next((ix[h] for h in hdr if h.endswith(w)), None)
            try:
                d[w] = float(r[j].replace(",", "")) if j is not None and r[j] not in ("", "n/a") else None
            except ValueError:
                d[w] = None
        d["name"] = nm
        per[cls].append(d)


def report(per, outdir, reps):
    traffic, lines = {}, []
    for cls, ks in per.items():
        tot = [(k["dram__bytes_read.sum"] or 0) + (k["dram__bytes_write.sum"] or 0) for k in ks]
        traffic[cls] = sum(tot) / len(tot)
        lines.append(f"== {cls}: {len(ks)} launches captured, mean DRAM traffic {traffic[cls] / 1e6:.1f} MB per launch")
        seen = defaultdict(list)
        for k in ks:
            seen[k["name"].split("(")[0][-60:]].append(k)
        for nm, group in seen.items():
            def avg(key):
                v = [g[key] for g in group if g[key] is not None]
                return sum(v) / len(v) if v else float("nan")
            lines.append(f"   {nm:60s} n={len(group):3d} time {avg('gpu__time_duration.sum') / 1e3:9.1f} us  dram "
                         f"{(avg('dram__bytes_read.sum') + avg('dram__bytes_write.sum')) / 1e6:8.1f} MB  tensor pipe "
                         f"{avg('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'):5.1f} %  tmem "
                         f"{avg('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):5.1f} %  dram "
                         f"{avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} %  issue "
                         f"{avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %  regs "
                         f"{avg('launch__registers_per_thread'):.0f}")
    (outdir / "r02_ncu_traffic.json").write_text(json.dumps(
        {"source": f"ncu --set full --clock-control none, {', '.join(Path(r).name for r in reps)}; dram__bytes_read.sum + dram__bytes_write.sum, "
                   "mean per launch of the class", "bytes_per_launch": traffic}, indent=1))
    (outdir / "r02_ncu_kernels.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
