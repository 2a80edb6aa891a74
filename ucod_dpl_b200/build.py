"""Build the C-ABI shared library `libucod_b200.so` in-tree with nvcc for sm_100a.

Each `csrc/*.cu` is compiled to an object (in parallel, skipped when up to date) and linked into
`ucod_dpl_b200/csrc/libucod_b200.so`.  No torch headers are involved: the library is plain CUDA C++
behind `include/ucod_b200.h`.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = CSRC / "build"
LIB_PATH = CSRC / "libucod_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    # IEEE-exact fp64 box maths and resample coefficients rely on explicit _rn intrinsics; fp32 FMA
    # contraction stays enabled for the float kernels.
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the ucod_b200 CUDA library cannot be built")
    return exe


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG_DIR.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hdrs), default=0.0)


def _compile_one(src: Path, force: bool, hdr_mtime: float) -> tuple[Path, str]:
    obj = OBJ_DIR / (src.stem + ".o")
    if (not force and obj.exists() and obj.stat().st_mtime >= src.stat().st_mtime
            and obj.stat().st_mtime >= hdr_mtime):
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("UCOD_NVCC_EXTRA", "").split(), "-c", str(src), "-o", str(obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    hdr_mtime = _deps_mtime()
    objs, logs = [], []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, log in ex.map(lambda s: _compile_one(s, force, hdr_mtime), srcs):
            objs.append(obj)
            if log:
                logs.append(log)
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    if verbose and logs:
        (OBJ_DIR / "ptxas.log").write_text("\n".join(logs))
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
