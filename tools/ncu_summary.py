#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + the hottest source lines by stall samples.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--lines N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max ", "launch__registers_per_thread ", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct",
        "gpu__dram_throughput.avg.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum ", "sm__throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:100])
        for h, u, v in zip(hdr, units, r):
            if any((h + " ").startswith(k) or k.strip() == h for k in KEYS) or "issue_stalled" in h and "ratio" in h:
                print(f"  {h:95s} {v} {u}")
    src = run([rep, "--page", "source", "--csv", "--print-source", "cuda"] if False else [rep, "--page", "source", "--csv"])
    rows = list(csv.reader(io.StringIO(src)))
    hi = next((i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r), None)
    if hi is None:
        print("(no source page)")
        return
    h = rows[hi]
    ci = h.index("Source")
    si = h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in rows[hi + 1:] if len(r) == len(h)]
    agg = {h[i]: sum(float(r[i] or 0) for r in body) for i in stall_cols}
    tot_s = sum(agg.values()) or 1.0
    print("== stall reasons:", ", ".join(f"{k[6:]} {v / tot_s * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    tot = sum(float(r[si] or 0) for r in body) or 1.0
    body.sort(key=lambda r: -float(r[si] or 0))
    print(f"== hottest lines by '{h[si]}' (total {tot:.0f})")
    for r in body[:nlines]:
        top = sorted(((float(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"  {float(r[si] or 0) / tot * 100:5.1f}%  {r[ci].strip()[:90]:90s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")


if __name__ == "__main__":
    main()
