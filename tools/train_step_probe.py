"""First-stage training step (BASELINE.json configs[4]: 16 cached key maps per GPU) — timing probe and ncu target."""
import sys
import time
from pathlib import Path
from types import SimpleNamespace
import torch
from safetensors.torch import load_file
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ucod_dpl_b200.models.discriminator import Discriminator
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.train import FirstStageTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
torch.manual_seed(31)
D = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=68)).cuda().train()
m = baseline(SimpleNamespace(dim=768))
m.load_state_dict(load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors")), strict=True)
tr = FirstStageTrainer(m.cuda().train(), D, lr0=2e-4, **({"use_graph": True} if graph else {}))
tr.cur_epoch = 3
g = torch.Generator().manual_seed(1)
tok = torch.randn(16, 1369, 768, generator=g).to(torch.bfloat16).cuda()
pl = (torch.rand(16, 1, 16, 16, generator=g) < 0.35).float().cuda()
for _ in range(10):
    tr.process_batch(tok, (37, 37), pl)
if graph and tr.graph_inputs() is not None:   # batch assembled in the graph's input buffers: no per-step copy
    bufs = tr.graph_inputs()
    bufs[0].copy_(tok), bufs[1].copy_(pl)
    tok, pl = bufs
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    loss = tr.process_batch(tok, (37, 37), pl)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"train step: {e0.elapsed_time(e1) / steps:.4f} ms device, {(t1 - t0) / steps * 1e3:.4f} ms host enqueue; loss {float(loss):.5f}")
