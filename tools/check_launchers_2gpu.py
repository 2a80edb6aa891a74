"""Multi-GPU check of the launchers on a synthetic image folder (run on a 2-GPU box):
   1. scripts.eval with 1 process and with 2 ranks (torchrun): identical result tables and PNG files;
   2. scripts.train with 2 ranks: runs the schedule, both ranks end with the same weights.
usage: python tools/check_launchers_2gpu.py [workdir]"""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
work = Path(sys.argv[1] if len(sys.argv) > 1 else "/tmp/ucod_launch_check").resolve()
work.mkdir(parents=True, exist_ok=True)
os.chdir(work)

import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("t", ROOT / "tests" / "test_data_eval_gpu.py")
t = importlib.util.module_from_spec(spec)
spec.loader.exec_module(t)

data = work / "data"
if not data.exists():
    t._write_set(data, "TE-X", 9, 0)
    t._write_set(data, "TR-X", 24, 40)
os.makedirs(work / "configs" / "uscod", exist_ok=True)
os.makedirs(work / "configs" / "__base__", exist_ok=True)
(work / "configs" / "uscod" / "tiny.py").write_text(
    (ROOT / "configs" / "uscod" / "UCOD-DPL_dinov2.py").read_text().replace("(518, 518)", "(224, 224)"))
(work / "configs" / "__base__" / "shared_defaults.py").write_text(
    (ROOT / "configs" / "__base__" / "shared_defaults.py").read_text()
    .replace("TR-CAMO+TR-COD10K", "TR-X").replace("TE-CAMO", "TE-X"))
ckpt = str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors")
env = dict(os.environ, PYTHONPATH=str(ROOT))
PRINT = "import json,sys; sys.path.insert(0, %r); from ucod_dpl_b200.scripts import %s as m; r = m.main(%r); print('RESULT', json.dumps(r, default=str))"


def run(nproc, module, argv, port):
    code = PRINT % (str(ROOT), module, argv)
    cmd = [sys.executable, "-c", code] if nproc == 1 else [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
        "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", code]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stdout[-2000:], p.stderr[-3000:])
        raise SystemExit(f"{module} with {nproc} process(es) failed")
    return [json.loads(l.split("RESULT ", 1)[1]) for l in p.stdout.splitlines() if l.startswith("RESULT ")]


def png_digest(run_dir):
    h = hashlib.sha256()
    for f in sorted((run_dir / "preds" / "TE-X").iterdir()):
        h.update(f.name.encode() + f.read_bytes())
    return h.hexdigest()


base = ["--config", "configs/uscod/tiny.py", "--work_dir", "work", "--load_from", ckpt, "--dataset_dir", str(data),
        "--datasets", "TE-X", "--batch_size", "4"]
r1 = run(1, "eval", base + ["--exp_name", "e1"], 0)[0]
r2 = run(2, "eval", base + ["--exp_name", "e2"], 29541)
assert len(r2) == 2 and r2[0] == r2[1], "ranks disagree on the reduced table"
for k, v in r1["TE-X"].items():
    assert abs(v - r2[0]["TE-X"][k]) < 1e-9, (k, v, r2[0]["TE-X"][k])
d1, d2 = png_digest(work / "work/uscod/tiny/e1"), png_digest(work / "work/uscod/tiny/e2")
assert d1 == d2, "PNG outputs differ between 1 and 2 ranks"
print("eval: 1 process == 2 ranks (table and PNGs)", r1["TE-X"])

tr = run(2, "train", ["--config", "configs/uscod/tiny.py", "--work_dir", "work", "--dataset_dir", str(data), "--cache_dir",
                      str(work / "cache"), "--max_epoch", "6", "--exp_name", "t2", "--no_save", "--batch_size", "4"], 29542)
assert len(tr) == 2 and tr[0]["best"] == tr[1]["best"], "ranks disagree after training"
assert tr[0]["weights_l1"] == tr[1]["weights_l1"], ("ranks ended with different weights", tr[0]["weights_l1"], tr[1]["weights_l1"])
ck = work / "work/uscod/tiny/t2/ckp/epoch5.pth/model.safetensors"
assert ck.exists()
print("train: 2 ranks ok, best", tr[0]["best"])
