#!/usr/bin/env python
"""Benchmark of the UCOD-DPL hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

One "step" = first-stage eval WITH Look-Twice of one synthetic batch of 64 images at 518x518 per GPU
(BASELINE.json metric "UCOD-DPL+LookTwice" on configs[1]; SURVEY.md 8(d) "LT" row, reference
engine/runner/loop_UCOD_DPL.py:297-352): first look (ViT key extraction -> DBA decoder -> upsample + binarise) ->
connected components / boxes -> crop + resize of 2 planted small objects per image -> second look (ViT + decoder on
the 128 crops) -> bicubic paste -> final bilinear resize + threshold.  Three backbone passes per image.
`value` = images/s with inputs resident in HBM, `e2e` = the same through the public pipeline call with pinned-host
inputs and a device->host read of the final masks.  The same JSON line carries the first-look-only rate, the other
BASELINE configurations (`workloads`: pseudo-labels 256 @224, training step 16 with the NCCL gradient all-reduce when
N > 1, CORAL 8 @1036) and per-kernel-class roofline fractions.
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
LOOKS = 2  # planted objects (= second looks) per image
WORKLOAD = ("UCOD-DPL_dinov2 first-stage eval with Look-Twice, synthetic batch 64 @518x518 per GPU, 2 planted small "
            "objects per image = 3 backbone passes per image (BASELINE.json configs[1], SURVEY.md 8(d) LT row)")
# algorithmic work per backbone pass, SURVEY.md §8(d): 11 full layers + last-layer LN/K-proj + patch embed
VIT_GFLOP_PER_PASS = 279.6
NAMES = ["gemm", "attention", "layernorm", "embed", "decoder", "resample", "pseudo_label", "ccl", "other"]
KC = len(NAMES)


def _traffic_table() -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel class, from the round's `ncu --set full`
    capture of this command (tools/ncu_traffic.py writes the file; None when no capture is committed)."""
    p = ROOT / "profiles" / "r02_ncu_traffic.json"
    try:
        return json.loads(p.read_text()).get("bytes_per_launch", {})
    except Exception:
        return {}


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
def _cpu_step_fn(sample: int):
    """The oracle port of the same step (HF-equivalent fp32 ViT + RevDecoder + process_preds + look_twice on the two
    planted boxes + final resize) on the host cores, all threads; returns (callable, cores)."""
    from oracle import pipeline as opipe
    from oracle import vit as ovit
    from ucod_dpl_b200.synth import planted_object_batch, synth_batch_u8
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = ovit.spec_for("dinov2")
    vit_sd = ovit.random_vit_state_dict(spec, seed=0)
    dec_sd = _load_decoder_sd()
    imgs = synth_batch_u8(0, sample, IMAGE, IMAGE)
    planted = planted_object_batch(0, sample, FEATURE, LOOKS)

    def step(n=sample):
        return opipe.look_twice_eval(vit_sd, spec, dec_sd, imgs[:n], (IMAGE, IMAGE), FEATURE, 0.15, "dynamic",
                                     first_logits=planted[:n])
    return step, cores


def run_reference(args, rank: int, world: int) -> None:
    """Reference arm: the reference algorithm for this path on the box's host cores, bounded sample per step."""
    if rank != 0:
        return
    sample = args.ref_batch
    step, cores = _cpu_step_fn(sample)
    step(1)  # warm-up (one image = 3 backbone passes)
    steps = max(1, min(args.steps, args.ref_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = steps * sample / dt
    line = {
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} images/step on host CPU"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {sample} images @518x518 (3 backbone passes each), torch CPU "
                                   "fp32 oracle port of the Look-Twice eval step"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args) -> dict:
    n, reps = max(2, args.ref_batch // 2), 2
    step, cores = _cpu_step_fn(n)
    step(1)
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = time.perf_counter() - t0
    return {"value": reps * n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {n} images @518x518 (3 backbone passes each), torch CPU fp32 oracle port of the same "
                      "Look-Twice eval step"}


# --------------------------------------------------------------------------------------------------
class _Prof:
    """per-kernel-class CUDA-event brackets of the library (two event records per launch: kept out of the headline
    region, used in a separate pass)."""

    def __init__(self, lib):
        self.lib = lib

    def run(self, fn, steps, dev):
        lib = self.lib
        lib.ucod_prof_collect(None, None, None)
        lib.ucod_prof_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize(dev)
        lib.ucod_prof_enable(0)
        ms_c, work_c, n_c = (ctypes.c_double * KC)(), (ctypes.c_double * KC)(), (ctypes.c_longlong * KC)()
        lib.ucod_prof_collect(ms_c, work_c, n_c)
        return e0.elapsed_time(e1), list(ms_c), list(work_c), list(n_c)


def lt_step_work(B: int, boxes, nbox, orig_hw) -> dict:
    """Algorithmic work of ONE Look-Twice eval step per kernel class (SURVEY.md 8(d) per-unit figures x the units of
    the step; FLOPs for the tensor classes, bytes for the rest).  The second pass runs with device-side counts, so
    the library cannot book its work on the host: it is stated here from the workload (B images, the step's boxes)."""
    T, P, D = 1370, 1369, 768
    n_look = int(sum(max(int(n), 0) for n in nbox))
    passes = B + n_look
    w = {}
    w["gemm"] = passes * (11 * 24.0 * T * D * D + 2.0 * T * D * D + 2.0 * P * 588 * D) + passes * 2.0 * P * D * 128
    w["attention"] = passes * 11 * 4.0 * T * T * D
    w["layernorm"] = passes * 23 * T * D * 6.0
    w["embed"] = passes * (3.0 * IMAGE * IMAGE + P * 640 * 2.0 + D * 4.0)
    w["decoder"] = passes * (P * D * 2.0 + 2 * 0) + B * FEATURE * FEATURE * 4.0 + n_look * 37 * 37 * 4.0
    crop = paste = 0.0
    for b in range(len(nbox)):
        for i in range(max(int(nbox[b]), 0)):
            x, y, bw, bh = [int(v) for v in boxes[b][i]]
            crop += 3.0 * (bw * orig_hw[1] / IMAGE) * (bh * orig_hw[0] / IMAGE) + 3.0 * IMAGE * IMAGE
            paste += 37 * 37 * 4.0 + bw * bh
    px = float(IMAGE * IMAGE)
    # first upsample (logits in, u8 mask out), canvas (mask -> x255), /255, final resize + threshold
    w["resample"] = B * (FEATURE * FEATURE * 4.0 + px) + B * 2 * px + B * 5 * px + B * 5 * px + crop + paste
    w["ccl"] = B * 9.0 * px
    return w


def _kernel_table(ms_c, work_c, n_c, steps, ms_total, peaks, work_override=None):
    """launches / time / share per kernel class plus the achieved rate against its roofline (tensor classes: FLOPs
    over the sustained bf16 peak; the rest: algorithmic bytes over the measured HBM copy bandwidth)."""
    kern = {}
    if work_override is not None:
        work_c = [work_override.get(n, 0.0) * steps for n in NAMES]
    for i, n in enumerate(NAMES):
        if not n_c[i]:
            continue
        k = {"launches_per_step": n_c[i] / steps, "ms_per_step": ms_c[i] / steps,
             "share": ms_c[i] / ms_total if ms_total else None}
        if ms_c[i] > 0:
            if i < 2:
                k["tflops"] = work_c[i] / (ms_c[i] * 1e-3) / 1e12
                k["frac_of_tensor_peak"] = k["tflops"] / peaks["tensor"]
            else:
                k["gbs"] = work_c[i] / (ms_c[i] * 1e-3) / 1e9
                k["frac_of_hbm_peak"] = k["gbs"] / peaks["hbm"]
        kern[n] = k
    return kern


def _timed(fn, steps, warmup, dev):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def _side_workloads(args, dev, lib, vit_sd, dec_sd, ext, model, peaks, rank, world, dist):
    """The other BASELINE.json configurations, each with its dominant kernel's roofline fraction.  Every rank runs
    them (data parallel, per-rank inputs); the training step all-reduces its decoder gradients over NCCL when
    world > 1, so the driver's scaling run exercises the path's only data-path collective."""
    from types import SimpleNamespace

    from ucod_dpl_b200.synth import random_refiner_state_dict, synth_batch_u8
    out = {}
    prof = _Prof(lib)

    def finish(name, n_img, ms, kern, extra=None):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dom = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
        entry = {"images_per_step_per_gpu": n_img, "ms_per_step": ms, "images_per_s": world * n_img / ms * 1e3,
                 "dominant_kernel": dom, "kernels": kern}
        if dom:
            entry["roofline_frac"] = kern[dom].get("frac_of_tensor_peak", kern[dom].get("frac_of_hbm_peak"))
        if extra:
            entry.update(extra)
        out[name] = entry

    # configs[2]: APM pseudo-label generation, 256 images @224 per launch
    from ucod_dpl_b200.generate_pseudo_label import PseudoLabelGenerator
    gen = PseudoLabelGenerator(vit_sd, "dinov2")
    imgs224 = synth_batch_u8(100000 + rank * 256, 256, 224, 224).to(dev)
    ms = _timed(lambda i: gen(imgs224), 5, 3, dev)
    tot, ms_c, work_c, n_c = prof.run(lambda i: gen(imgs224), 3, dev)
    finish("pseudo_label_256x224", 256, ms, _kernel_table(ms_c, work_c, n_c, 3, tot, peaks))
    del gen, imgs224

    # configs[4]: first-stage training step, 16 cached key maps per GPU, gradient all-reduce over NCCL
    from ucod_dpl_b200.models.discriminator import Discriminator
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.train import FirstStageTrainer
    torch.manual_seed(31)
    D = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=FEATURE)).to(dev).train()
    m2 = baseline(SimpleNamespace(dim=768))
    m2.load_state_dict(dec_sd, strict=True)
    tr = FirstStageTrainer(m2.to(dev).train(), D, lr0=2e-4, use_graph=True)   # CUDA-graph replay of fwd / APM / bwd
    tr.cur_epoch = 3
    g = torch.Generator().manual_seed(1 + rank)
    tok = torch.randn(16, 1369, 768, generator=g).to(torch.bfloat16).to(dev)
    pl = (torch.rand(16, 1, 16, 16, generator=g) < 0.35).float().to(dev)
    for _ in range(4):                      # two eager steps, the capture, one replay
        tr.process_batch(tok, (37, 37), pl)
    bufs = tr.graph_inputs()                # the batch lives in the graph's own input buffers, as a launcher that
    if bufs is not None:                    # index_selects its HBM-resident training set into them would leave it
        bufs[0].copy_(tok), bufs[1].copy_(pl)
        tok, pl = bufs
    ms = _timed(lambda i: tr.process_batch(tok, (37, 37), pl), 100, 20, dev)
    tot, ms_c, work_c, n_c = prof.run(lambda i: tr.process_batch(tok, (37, 37), pl), 20, dev)
    finish("train_step_16", 16, ms, _kernel_table(ms_c, work_c, n_c, 20, tot, peaks),
           {"collective": f"NCCL all-reduce of {tr.n} fp32 decoder gradients per step" if world > 1 else "none (1 GPU)",
            "cuda_graph": True})
    del tr, tok, pl

    # configs[3]: CORAL second-stage eval, 8 originals @1036^2 per launch (80 backbone passes + refiner windows)
    if not args.no_coral:
        from ucod_dpl_b200.engine.runner.loop_CORAL import CoralEvaluator
        from ucod_dpl_b200.models.UDLR import SparseRefiner
        ref = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015))
        ref.load_state_dict(random_refiner_state_dict(0), strict=True)
        ev = CoralEvaluator(ext, model, ref.to(dev).eval(), (IMAGE, IMAGE), 3, 56)
        big = synth_batch_u8(200000 + rank * 8, 8, 1036, 1036).to(dev)
        ms = _timed(lambda i: ev(big), 2, 1, dev)
        tot, ms_c, work_c, n_c = prof.run(lambda i: ev(big), 1, dev)
        finish("coral_eval_8x1036", 8, ms, _kernel_table(ms_c, work_c, n_c, 1, tot, peaks))
        del ev, big
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    from types import SimpleNamespace

    from ucod_dpl_b200 import _lib, ops
    from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import LookTwiceEvaluator
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.synth import planted_object_batch, random_vit_state_dict, synth_batch_u8
    from ucod_dpl_b200.vit import VitKeyExtractor, spec_for

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
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    dec_sd = _load_decoder_sd()
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(dec_sd, strict=True)
    model = model.to(dev).eval()
    ext = VitKeyExtractor(vit_sd, spec_for("dinov2"), device=dev)
    ev = LookTwiceEvaluator(ext, model, (IMAGE, IMAGE), FEATURE, 0.15, "dynamic",
                            max_looks_per_image=args.max_looks)

    B, NB = args.batch, args.rotate
    host = [synth_batch_u8((rank * NB + i) * B, B, IMAGE, IMAGE).pin_memory() for i in range(NB)]
    dev_in = [h.to(dev) for h in host]
    planted = [planted_object_batch((rank * NB + i) * B, B, FEATURE, LOOKS).to(dev) for i in range(NB)]

    def step(images, logits):
        """loop_UCOD_DPL.py:297-317 for one batch: -> (uint8 final masks [B,S,S], LookTwiceResult)."""
        res = ev.look_twice_device(images, first_logits=logits)
        return ops.upsample_bilinear(res.final, (IMAGE, IMAGE), binarize=3), res

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    last = None
    for i in range(args.warmup):
        last = step(dev_in[i % NB], planted[i % NB])
        last[1].check()  # outside the timed regions: also lets the evaluator settle on the chunks it enqueues ahead
    barrier()
    # the planted objects must produce exactly LOOKS second looks per image
    kept, status, wanted, _ = last[1].counts.cpu().tolist()
    if kept != LOOKS * B or status:
        raise SystemExit(f"benchmark set-up error: {kept} second looks for {B} images (status {status})")
    step_work = lt_step_work(B, last[1].boxes.cpu().tolist(), last[1].nbox.cpu().tolist(), (IMAGE, IMAGE))

    # ---- timed region 1: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.ucod_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(dev_in[i % NB], planted[i % NB])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.ucod_launch_count() - launches0

    # ---- per-kernel-class breakdown of the same K steps (separate pass, see _Prof) ----
    peaks = _peaks()
    ms_prof_total, ms_c, work_c, n_c = _Prof(lib).run(lambda i: step(dev_in[i % NB], planted[i % NB]), args.steps, dev)

    # ---- secondary: the first look alone (first-stage eval without Look-Twice, round 1's headline) ----
    ev_first = LookTwiceEvaluator(ext, model, (IMAGE, IMAGE), FEATURE, 0.15, "dynamic", look_twice=False)
    barrier()
    ms_first = _timed(lambda i: ops.upsample_bilinear(ev_first.first_look(dev_in[i % NB])[:, 0], (IMAGE, IMAGE),
                                                      binarize=True), max(3, args.steps // 2), 1, dev)

    # ---- timed region 2: end to end through the public call, pinned host in, final masks read back to the host ----
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
            masks, _ = step(dev_buf[cur], planted[i % NB])
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

    t = torch.tensor([ms_total, ms_e2e, ms_first], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_first = t.tolist()

    del dev_buf, out_dev, host, dev_in
    torch.cuda.empty_cache()
    workloads = None
    if not args.no_workloads:
        workloads = _side_workloads(args, dev, lib, vit_sd, dec_sd, ext, model, peaks, rank, world, dist)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    kern = _kernel_table(ms_c, work_c, n_c, args.steps, ms_prof_total, peaks, work_override=step_work)
    traffic = _traffic_table()
    dom = max(range(KC), key=lambda i: ms_c[i])
    dk = kern[NAMES[dom]]
    if dom in (0, 1):
        roof = {"kernel": NAMES[dom], "bound": "tensor", "achieved": dk["tflops"], "peak": peaks["tensor"],
                "unit": "TFLOP/s", "frac": dk["frac_of_tensor_peak"],
                "traffic": traffic.get(NAMES[dom]) if B == 64 else None,
                "peak_source": peaks["source"] + ", sustained bf16",
                "traffic_source": "profiles/r02_ncu_traffic.json (ncu --set full, mean per launch)"}
    else:
        roof = {"kernel": NAMES[dom], "bound": "hbm", "achieved": dk["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                "frac": dk["frac_of_hbm_peak"], "traffic": traffic.get(NAMES[dom]) if B == 64 else None,
                "peak_source": peaks["source"]}
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)
    passes = 1 + LOOKS
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "image": IMAGE, "feature_size": FEATURE,
                   "second_looks_per_image": LOOKS, "backbone_passes_per_image": passes,
                   "second_look_capacity_per_image": args.max_looks,
                   "weights": "random-init ViT-B/14 (seed 0) + weights/UCOD_DPL_dinov2.safetensors",
                   "l2": f"inputs rotate over {NB} batches ({NB * B * 3 * IMAGE * IMAGE / 1e6:.0f} MB) and the "
                         "per-step activation working set (~1.5 GB) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": B * 3 * IMAGE * IMAGE,
                "d2h_bytes_per_step": B * IMAGE * IMAGE, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "host_syncs_per_step": 0,
        "second_look_chunks_enqueued": min(args.max_looks, max(ev._recent_chunks or [args.max_looks])),
        "roofline": roof,
        "kernels": kern,
        "vit_tensor_frac": (passes * VIT_GFLOP_PER_PASS * 1e9 * B / (ms_step * 1e-3) / 1e12) / peaks["tensor"],
        "first_look_only": {"images_per_s": world * B / ms_first * 1e3, "ms_per_step": ms_first,
                            "vit_tensor_frac": (VIT_GFLOP_PER_PASS * 1e9 * B / (ms_first * 1e-3) / 1e12) / peaks["tensor"],
                            "what": "first-stage eval without Look-Twice (ViT -> decoder -> upsample + binarise)"},
    }
    if workloads is not None:
        line["workloads"] = workloads
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--rotate", type=int, default=4, help="number of distinct input batches cycled through")
    ap.add_argument("--max-looks", type=int, default=4, help="second-look capacity per image (chunks of one batch)")
    ap.add_argument("--ref-batch", type=int, default=4, help="images per step of the CPU reference arm")
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the pseudo-label / training / CORAL block")
    ap.add_argument("--no-coral", action="store_true")
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
