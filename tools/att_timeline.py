"""Debug: per-tile clock64 timeline of one attention CTA (library built with UCOD_NVCC_EXTRA=-DUCOD_ATT_TIMELINE)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib
if len(sys.argv) > 1:
    _lib._LIB_PATH = Path(sys.argv[1]).resolve()
import os
B, H, T, D = int(os.environ.get('ATT_B', 64)), 12, int(os.environ.get('ATT_T', 1370)), 64
qkv = torch.randn(B, T, 3 * H * D, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
ld = 3 * H * D
args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D, B, H, D,
        T, T, _lib.c_float(0.125), _lib.stream_ptr())
for _ in range(3):
    _lib.call("ucod_attention", *args)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (6 * 16 * 8))()
_lib.load().ucod_debug_att_timeline(buf)
g = lambda role, j, s: buf[(role * 16 + j) * 8 + s]
print("issuers per tile j: [kfull(j), sfree(j-1), S(j) issued, vfull(j), p(j), PV(j) issued] | S issue, wait p, PV issue")
NT = (T + 127) // 128
for j in range(NT):
    t = [g(1, j, s) for s in range(6)]
    print("  tile", j, t, "|", t[2] - t[1], t[4] - t[3], t[5] - t[4])
for w in range(4):
    print(f"softmax warp {4 + w}: start | wait_s, ld, max, exp, wait_pv, st, arrive | period")
    for j in range(0, NT - 1):
        t = [g(2 + w, j, s) for s in range(8)]
        nxt = g(2 + w, j + 1, 0)
        print("  tile", j, t[0], "|", [t[1] - t[0], t[2] - t[1], t[3] - t[2], t[7] - t[3], (t[4] - t[7]) if j else 0, t[5] - (t[4] if j else t[7]), t[6] - t[5]], "|", nxt - t[0])
