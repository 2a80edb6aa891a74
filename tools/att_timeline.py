"""Debug: per-tile clock64 timeline of one attention CTA (library built with UCOD_NVCC_EXTRA=-DUCOD_ATT_TIMELINE)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib
B, H, T, D = 64, 12, 1370, 64
qkv = torch.randn(B, T, 3 * H * D, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
ld = 3 * H * D
args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D, B, H, D,
        T, T, _lib.c_float(0.125), _lib.stream_ptr())
for _ in range(3):
    _lib.call("ucod_attention", *args)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 16 * 8))()
_lib.load().ucod_debug_att_timeline(buf)
names = {0: ["kempty", "vempty"], 1: ["kfull", "sfree", "S issued", "vfull", "p", "PV issued"],
         2: ["start", "s ready", "ld done", "max done", "pv ready", "exp done", "p arrived"]}
for role, rn in ((0, "producer"), (1, "mma"), (2, "softmax")):
    print(rn, names[role])
    for j in range(11):
        print("  tile", j, [buf[(role * 16 + j) * 8 + s] for s in range(len(names[role]))])
