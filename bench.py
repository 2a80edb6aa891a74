#!/usr/bin/env python
"""Benchmark of the UCOD-DPL hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

One "step" = first-stage eval (ViT key extraction -> DBA decoder -> upsample + binarise) of one synthetic batch
of 64 images at 518x518 per GPU (BASELINE.json configs[1]).  `value` = images/s with inputs resident in HBM,
`e2e` = the same through the public pipeline call with pinned-host inputs and a device->host read of the masks.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

IMAGE = 518
FEATURE = 68
WORKLOAD = "UCOD-DPL_dinov2 first-stage eval, synthetic batch 64 @518x518 per GPU (BASELINE.json configs[1])"
# algorithmic work per image, SURVEY.md §8(d): 11 full layers + last-layer LN/K-proj + patch embed
VIT_GFLOP_PER_IMAGE = 279.6


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class from the round's
# `ncu --set full` capture of this very command at batch 64 (profiles/r01_gemm2_ncu_summary.txt: mean over the 47
# GEMM launches of a step; per shape 488 / 630 / 616 / 1058 MB, each <= the algorithmic bytes).  None = no capture.
NCU_TRAFFIC_BYTES = {"gemm": 6.73e8, "attention": 5.21e8}  # attention: profiles/r01_attention_final_ncu_summary.txt


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tensor": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "tensor_burst": float(d.get("bf16_tflops", 1590.0)), "hbm": float(d.get("hbm_gbs", 6650.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _load_decoder_sd():
    from safetensors.torch import load_file
    return load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))


# --------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int) -> None:
    """Reference arm: the reference algorithm for this path (oracle port of HF ViT + RevDecoder + process_preds
    upsample/threshold) on the box's host cores, all threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import pipeline as opipe
    from oracle import vit as ovit
    from ucod_dpl_b200.synth import synth_batch_u8
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = ovit.spec_for("dinov2")
    vit_sd = ovit.random_vit_state_dict(spec, seed=0)
    dec_sd = _load_decoder_sd()
    sample = args.ref_batch
    imgs = synth_batch_u8(0, sample, IMAGE, IMAGE)
    for _ in range(max(1, min(args.warmup, 1))):
        opipe.first_stage_eval(vit_sd, spec, dec_sd, imgs[:1], (IMAGE, IMAGE), FEATURE)
    steps = max(1, min(args.steps, args.ref_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        opipe.first_stage_eval(vit_sd, spec, dec_sd, imgs, (IMAGE, IMAGE), FEATURE)
    dt = time.perf_counter() - t0
    val = steps * sample / dt
    line = {
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} images/step on host CPU"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {sample} images @518x518, torch CPU fp32 oracle port"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    from types import SimpleNamespace

    from ucod_dpl_b200 import _lib
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.pipeline import FirstStageEval
    from ucod_dpl_b200.synth import synth_batch_u8
    from ucod_dpl_b200.vit import spec_for

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # weights: random-init ViT-B/14 (no DINOv2 weights offline), shipped decoder checkpoint
    from ucod_dpl_b200.synth import random_vit_state_dict
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(_load_decoder_sd(), strict=True)
    pipe = FirstStageEval(vit_sd, spec_for("dinov2"), model, (IMAGE, IMAGE), FEATURE, device=dev)

    B, NB = args.batch, args.rotate
    host = [synth_batch_u8((rank * NB + i) * B, B, IMAGE, IMAGE).pin_memory() for i in range(NB)]
    dev_in = [h.to(dev) for h in host]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        pipe(dev_in[i % NB])
    barrier()

    # ---- timed region 1: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.ucod_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        pipe(dev_in[i % NB])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.ucod_launch_count() - launches0

    # ---- per-kernel-class breakdown: the same K steps again with the library's CUDA-event brackets enabled (two
    # event records per launch cost ~1 % of a step, so they stay out of the headline region above) ----
    lib.ucod_prof_collect(None, None, None)
    lib.ucod_prof_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for i in range(args.steps):
        pipe(dev_in[i % NB])
    p1.record()
    barrier()
    lib.ucod_prof_enable(0)
    ms_prof_total = p0.elapsed_time(p1)
    KC = 9
    ms_c, work_c, n_c = (ctypes.c_double * KC)(), (ctypes.c_double * KC)(), (ctypes.c_longlong * KC)()
    lib.ucod_prof_collect(ms_c, work_c, n_c)

    # ---- timed region 2: end to end through the public call, pinned host in, masks read back to the host ----
    # Every step uploads its own input batch from pinned host memory and downloads its own masks; uploads of step
    # i+1 and downloads of step i-1 run on a copy stream while step i computes (double-buffered device inputs and
    # mask buffers), as a serving loop would.  All copies are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    dev_buf = [torch.empty_like(dev_in[0]) for _ in range(2)]
    out_dev = [torch.empty(B, IMAGE, IMAGE, dtype=torch.uint8, device=dev) for _ in range(2)]
    out_hosts = [torch.empty(B, IMAGE, IMAGE, dtype=torch.uint8).pin_memory() for _ in range(2)]
    up_done = [torch.cuda.Event() for _ in range(2)]
    comp_done = [torch.cuda.Event() for _ in range(2)]
    down_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        with torch.cuda.stream(copy_stream):
            dev_buf[0].copy_(host[0], non_blocking=True)
            up_done[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(comp_done[nxt])   # step i-1 no longer reads dev_buf[nxt]
                    dev_buf[nxt].copy_(host[(i + 1) % NB], non_blocking=True)
                    up_done[nxt].record(copy_stream)
            main.wait_event(up_done[cur])
            if i >= 2:
                main.wait_event(down_done[cur])                  # out_dev[cur] has been downloaded
            masks = pipe(dev_buf[cur])
            out_dev[cur].copy_(masks)
            comp_done[cur].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(comp_done[cur])
                out_hosts[cur].copy_(out_dev[cur], non_blocking=True)
                down_done[cur].record(copy_stream)
            if i >= 1:
                down_done[cur ^ 1].synchronize()                 # the host consumes step i-1's masks
        down_done[(n - 1) & 1].synchronize()

    e2e_loop(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_loop(args.steps)
    copy_stream.synchronize()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    names = ["gemm", "attention", "layernorm", "embed", "decoder", "resample", "pseudo_label", "ccl", "other"]
    kern = {}
    for i, n in enumerate(names):
        if n_c[i]:
            kern[n] = {"launches_per_step": n_c[i] / args.steps, "ms_per_step": ms_c[i] / args.steps,
                       "share": ms_c[i] / ms_prof_total if ms_prof_total else None}
    # dominant kernel class -> roofline entry (tensor classes: FLOPs; others: algorithmic bytes)
    dom = max(range(KC), key=lambda i: ms_c[i])
    if dom in (0, 1):
        achieved = work_c[dom] / (ms_c[dom] * 1e-3) / 1e12
        roof = {"kernel": names[dom], "bound": "tensor", "achieved": achieved, "peak": peaks["tensor"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
                "traffic": NCU_TRAFFIC_BYTES.get(names[dom]) if B == 64 else None,
                "peak_source": peaks["source"] + ", sustained bf16"}
    else:
        achieved = work_c[dom] / (ms_c[dom] * 1e-3) / 1e9
        roof = {"kernel": names[dom], "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": achieved / peaks["hbm"], "traffic": None, "peak_source": peaks["source"]}
    for i in (0, 1):
        if n_c[i]:
            kern[names[i]]["tflops"] = work_c[i] / (ms_c[i] * 1e-3) / 1e12
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "image": IMAGE, "feature_size": FEATURE,
                   "weights": "random-init ViT-B/14 (seed 0) + weights/UCOD_DPL_dinov2.safetensors",
                   "l2": f"inputs rotate over {NB} batches ({NB * B * 3 * IMAGE * IMAGE / 1e6:.0f} MB) and the "
                         "per-step activation working set (~1.5 GB) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": B * 3 * IMAGE * IMAGE,
                "d2h_bytes_per_step": B * IMAGE * IMAGE, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": kern,
        "vit_tensor_frac": (VIT_GFLOP_PER_IMAGE * 1e9 * B / (ms_step * 1e-3) / 1e12) / peaks["tensor"],
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args) -> dict:
    """The oracle port of the same step on the host cores, bounded sample (reported baseline, not the target)."""
    from oracle import pipeline as opipe
    from oracle import vit as ovit
    from ucod_dpl_b200.synth import synth_batch_u8
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = ovit.spec_for("dinov2")
    vit_sd = ovit.random_vit_state_dict(spec, seed=0)
    dec_sd = _load_decoder_sd()
    n, reps = args.ref_batch, 3
    imgs = synth_batch_u8(0, n, IMAGE, IMAGE)
    opipe.first_stage_eval(vit_sd, spec, dec_sd, imgs[:1], (IMAGE, IMAGE), FEATURE)
    t0 = time.perf_counter()
    for _ in range(reps):
        opipe.first_stage_eval(vit_sd, spec, dec_sd, imgs, (IMAGE, IMAGE), FEATURE)
    dt = time.perf_counter() - t0
    return {"value": reps * n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {n} images @518x518, torch CPU fp32 oracle port of the same step"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--rotate", type=int, default=4, help="number of distinct input batches cycled through")
    ap.add_argument("--ref-batch", type=int, default=8, help="images per step of the CPU reference arm")
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
