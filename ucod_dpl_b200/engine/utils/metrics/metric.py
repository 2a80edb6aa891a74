"""`statistics` — the COD metric suite with the reference's interface (engine/utils/metrics/metric.py:19-74:
`reset()`, `step(gt_tensor, pred_tensor)`, `get_result()` -> dict with ACC, mIOU, E_MAX, E_MEAN, F_MAX, F_MEAN,
SMeasure, MAE, WFM), computed per image in fp64 on the GPU (csrc/metrics.cu) instead of numpy/scipy on the CPU.
Per-image results stay on the device; `get_result` reduces them (and, under torch.distributed, across ranks with one
all-reduce of sums + count)."""
from __future__ import annotations

import ctypes

import torch

from .... import _lib
from ...._lib import ptr, stream_ptr

OUT = 519
_KEYS = ("acc", "iou", "mae", "sm", "em_adp", "fm_adp", "wfm")


def cod_metrics(gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """gt, pred: CUDA tensors [B,h,w] (any float / integer dtype) -> fp64 [B, 519] per-image measures."""
    _lib.require_cuda(gt, pred)
    g = gt.to(torch.float32).contiguous()
    p = pred.to(torch.float32).contiguous()
    if g.shape != p.shape or g.dim() != 3:
        raise _lib.UcodError(f"cod_metrics expects equally shaped [B,h,w] tensors, got {tuple(g.shape)} / {tuple(p.shape)}")
    B, h, w = g.shape
    dev = g.device
    out = torch.empty(B, OUT, device=dev, dtype=torch.float64)
    lib = _lib.load()
    lib.ucod_cod_metrics_workspace_bytes.restype = ctypes.c_uint64
    ws = torch.empty(int(lib.ucod_cod_metrics_workspace_bytes(B, h, w)) + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    with torch.cuda.device(dev):
        _lib.call("ucod_cod_metrics", ptr(g), ptr(p), B, h, w, ptr(out), ctypes.c_void_p(ws.data_ptr() + off),
                  ctypes.c_uint64(ws.numel() - off), stream_ptr(dev))
    return out


class statistics:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.reset()

    def reset(self):
        self._rows = []

    def step(self, gt_tensor: torch.Tensor, pred_tensor: torch.Tensor):
        """gt_tensor / pred_tensor: [B,h,w] or [B,1,h,w] (the eval loops pass the label and the binarised mask)."""
        g = gt_tensor.to(self.device)
        p = pred_tensor.to(self.device)
        if g.dim() == 4:
            g = g[:, 0]
        if p.dim() == 4:
            p = p[:, 0]
        self._rows.append(cod_metrics(g, p))

    def per_image(self) -> torch.Tensor:
        return torch.cat(self._rows, 0) if self._rows else torch.zeros(0, OUT, device=self.device, dtype=torch.float64)

    def get_result(self) -> dict:
        rows = self.per_image()
        from .... import dist as ud
        sums, count = ud.reduce_metric_sums(rows.sum(0), rows.shape[0])
        mean = (sums / max(count, 1)).cpu()
        em, fm = mean[7:7 + 256], mean[7 + 256:]
        r = {k: float(mean[i]) for i, k in enumerate(_KEYS)}
        return {"ACC": r["acc"], "mIOU": r["iou"], "E_MAX": float(em.max()), "E_MEAN": float(em.mean()),
                "F_MAX": float(fm.max()), "F_MEAN": float(fm.mean()), "SMeasure": r["sm"], "MAE": r["mae"],
                "WFM": r["wfm"]}
