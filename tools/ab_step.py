"""A/B builds of the library on the first-look eval step (ViT -> decoder -> mask, 64 images @518^2): step time and the
per-kernel-class split under the power cap, alternating subprocess runs.  usage: python tools/ab_step.py ab/lib_a.so ..."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import ctypes, sys, torch
sys.path.insert(0, %r)
from types import SimpleNamespace
from safetensors.torch import load_file
from ucod_dpl_b200 import _lib
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.pipeline import FirstStageEval
from ucod_dpl_b200.synth import random_vit_state_dict, synth_batch_u8
from ucod_dpl_b200.vit import spec_for
lib = _lib.load()
model = baseline(SimpleNamespace(dim=768)); model.load_state_dict(load_file(%r), strict=True)
pipe = FirstStageEval(random_vit_state_dict(spec_for("dinov2"), seed=0), spec_for("dinov2"), model, (518, 518), 68)
imgs = [synth_batch_u8(i * 64, 64, 518, 518).cuda() for i in range(2)]
for i in range(4): pipe(imgs[i & 1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(12): pipe(imgs[i & 1])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 12
lib.ucod_prof_collect(None, None, None); lib.ucod_prof_enable(1)
for i in range(6): pipe(imgs[i & 1])
torch.cuda.synchronize(); lib.ucod_prof_enable(0)
m, w, n = (ctypes.c_double * 9)(), (ctypes.c_double * 9)(), (ctypes.c_longlong * 9)()
lib.ucod_prof_collect(m, w, n)
print("step %%.2f ms | gemm %%.2f attn %%.2f ln %%.2f other %%.2f" %% (ms, m[0] / 6, m[1] / 6, m[2] / 6, sum(m[3:]) / 6))
''' % (str(ROOT), str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))
for rep in range(2):
    for p in sys.argv[1:]:
        env = dict(os.environ, UCOD_B200_LIB=str(Path(p).resolve()))
        out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env)
        print(f"{p:28s}", out.stdout.strip() if out.returncode == 0 else out.stderr[-400:], flush=True)
