"""PNG output of predicted masks (reference: engine/utils/save_image.py:6-73).  Same entry point and naming rules:
a 4-D batch with more than one mask goes to `{save_path without extension}/{i}.png`, a single mask to `save_path`
with `.jpg` replaced by `.png`; values are `mask * 255` as 8-bit greyscale.  Encoding is host work (PIL)."""
from __future__ import annotations

import os

import numpy as np
import torch


def _to_u8(mask: torch.Tensor) -> np.ndarray:
    return (mask.detach().to("cpu").numpy() * 255).astype(np.uint8)


def _write_png(arr: np.ndarray, path: str) -> None:
    from PIL import Image

    Image.fromarray(arr, mode="L").save(path)


def save_tensor_binary_mask_as_image(binary_mask: torch.Tensor, save_path: str) -> None:
    try:
        if binary_mask.dim() == 4 and binary_mask.shape[0] > 1:
            folder = os.path.splitext(save_path)[0]
            os.makedirs(folder, exist_ok=True)
            for i in range(binary_mask.shape[0]):
                m = binary_mask[i].squeeze()
                if m.dim() != 2:
                    print(f"Warning: Unexpected mask dimensions after squeeze: {tuple(m.shape)} for batch item {i}")
                    continue
                _write_png(_to_u8(m), os.path.join(folder, f"{i}.png"))
            return
        m = binary_mask.squeeze()
        if m.dim() != 2:
            print(f"Warning: Could not save mask due to unexpected shape: {tuple(binary_mask.shape)}")
            return
        os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
        _write_png(_to_u8(m), save_path.replace(".jpg", ".png"))
    except Exception as e:  # the reference reports and carries on
        print(f"Error saving mask to {save_path}: {e}")


class AsyncMaskWriter:
    """PNG encoding off the critical path: `submit(mask, path)` copies the mask to the host and hands the encode to a
    thread pool (zlib releases the GIL); `close()` waits for the files.  Same naming rules as
    `save_tensor_binary_mask_as_image`."""

    def __init__(self, workers: int = 8):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(workers)
        self._pending = []

    def submit(self, binary_mask: torch.Tensor, save_path: str) -> None:
        host = binary_mask.detach().to("cpu")
        self._pending.append(self._pool.submit(save_tensor_binary_mask_as_image, host, save_path))

    def close(self) -> None:
        for f in self._pending:
            f.result()
        self._pending.clear()
        self._pool.shutdown()
