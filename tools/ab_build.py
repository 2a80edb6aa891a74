"""Build variants of the library for same-box A/B runs: every variant recompiles ONE source with extra -D flags and
links it with the up-to-date objects of the regular build.
usage: python tools/ab_build.py attention.cu name1="-DX=1 -DY=2" name2="..."   ->  ab/lib_<name>.so"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ucod_dpl_b200 import build as b  # noqa: E402

b.build()
src = b.CSRC / sys.argv[1]
out = ROOT / "ab"
out.mkdir(exist_ok=True)
others = [str(b.OBJ_DIR / (c.stem + ".o")) for c in sorted(b.CSRC.glob("*.cu")) if c.stem != src.stem]
for spec in sys.argv[2:]:
    name, flags = spec.split("=", 1)
    obj = out / f"{src.stem}_{name}.o"
    cmd = [b._nvcc(), *b.NVCC_FLAGS, *flags.split(), "-c", str(src), "-o", str(obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise SystemExit(p.stderr)
    regs = [ln for ln in p.stderr.splitlines() if "registers" in ln or "spill" in ln]
    lib = out / f"lib_{name}.so"
    subprocess.run([b._nvcc(), "-shared", "-o", str(lib), str(obj), *others, "-gencode",
                    "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
    print(lib, "|", flags)
    for r in regs:
        if "spill" in r and " 0 bytes spill stores" in r:
            continue
        print("    ", r.strip()[:160])
