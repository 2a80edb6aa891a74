"""ctypes binding of `libucod_b200.so` (the C-ABI declared in include/ucod_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_LIB_PATH = Path(os.environ.get("UCOD_B200_LIB") or Path(__file__).resolve().parent / "csrc" / "libucod_b200.so")  # env: A/B builds
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_double = ctypes.c_double
c_i64 = ctypes.c_int64


class UcodError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library (building is the job of `__graft_entry__.build()` / `ucod_dpl_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise UcodError(
            f"{_LIB_PATH} not found: the ucod_b200 CUDA library has not been built "
            "(run `python -m ucod_dpl_b200.build`); there is no CPU fallback.")
    lib = ctypes.CDLL(str(_LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else ctypes.DEFAULT_MODE)
    lib.ucod_last_error.restype = ctypes.c_char_p
    lib.ucod_last_error.argtypes = []
    lib.ucod_abi_version.restype = c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ucod_last_error().decode("utf-8", "replace")
        raise UcodError(f"{what or 'ucod call'} failed (status {rc}): {msg}")


def ptr(t) -> c_void_p:
    """Device (or host) pointer of a tensor, NULL for None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device=None) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise UcodError("ucod_b200 kernels need CUDA tensors; there is no CPU fallback")


def call(name: str, *args) -> None:
    """Call an `int`-returning C-ABI function and raise on failure."""
    fn = getattr(load(), name)
    fn.restype = c_int
    check(fn(*args), name)
