"""CORAL second-stage evaluation (reference: engine/runner/loop_CORAL.py:41-341 `LocalRefineValidationLoop` and the
feature production of data/datasets/lr_dataset.py:82-168), batched on the device.

`CoralEvaluator(images)` = for every image: low-res keys `l` (whole image @S), high-res keys `h` (3x3 windows of the
3x-upscaled image, each @S), both resampled to window_length^2, coarse logits from the first-stage decoder
(or the 4-patch 102^2 stitch of the m-path), optional centre-crop retry, SparseRefiner, `sigmoid -> bilinear -> >0.5`.
The reference runs this at batch 1 with one ViT call per window (10-11 launches per image); here all windows of
all images of a batch go through the backbone in one call.
"""
from __future__ import annotations

import torch

from ... import ops
from ...vit import VitKeyExtractor


class CoralEvaluator:
    def __init__(self, extractor: VitKeyExtractor, model, refiner, image_size, window_size: int = 3,
                 window_length: int = 56, require_m_patches: bool = False, vit_chunk: int = 64):
        self.extractor, self.model, self.refiner = extractor, model, refiner
        self.S = int(image_size[0])
        self.ws, self.g = int(window_size), int(window_length)
        self.patch = extractor.spec.patch
        self.require_m = bool(require_m_patches)
        self.vit_chunk = int(vit_chunk)

    # ---- reference-compatible helpers (loop_CORAL.py:62-96,168-204,247-258,313-341) ----
    @staticmethod
    def concate_preds(preds: torch.Tensor) -> torch.Tensor:
        """[b,4,c,68,68] -> [b,c,102,102]; the overlapping 2x2 patches are averaged."""
        b, n, c, h, w = preds.shape
        full = torch.zeros(b, c, 102, 102, device=preds.device)
        counter = torch.zeros(b, c, 102, 102, device=preds.device)
        for i in range(2):
            for j in range(2):
                full[:, :, i * 34:i * 34 + 68, j * 34:j * 34 + 68] += preds[:, i * 2 + j]
                counter[:, :, i * 34:i * 34 + 68, j * 34:j * 34 + 68] += 1.0
        return full / (counter + 1e-6)

    @staticmethod
    def _should_crop_center(preds: torch.Tensor) -> torch.Tensor:
        """per image: fraction of positive coarse logits < 0.001 (the reference evaluates this at batch 1)."""
        return (preds > 0).flatten(1).sum(1).float() / (preds.shape[2] * preds.shape[3]) < 0.001

    @staticmethod
    def _center_pad(x: torch.Tensor, fill_value: float = -10.0) -> torch.Tensor:
        b, c, h, w = x.shape
        out = torch.full((b, c, 2 * h, 2 * w), fill_value, device=x.device, dtype=x.dtype)
        out[:, :, h // 2:h // 2 + h, w // 2:w // 2 + w] = x
        return out

    @staticmethod
    def process_preds(preds: torch.Tensor, size) -> torch.Tensor:
        """refined logits (or probabilities) [B,1,s,s] -> uint8 masks [B,h,w] (`sigmoid -> bilinear -> > 0.5`)."""
        probs = bool(torch.all((preds >= 0) & (preds <= 1)))
        return ops.upsample_bilinear(preds[:, 0], size, binarize=3 if probs else 2)

    # ---- feature production (lr_dataset.py:82-168) ----
    def _keys(self, images_u8: torch.Tensor) -> torch.Tensor:
        out = []
        for i in range(0, images_u8.shape[0], self.vit_chunk):
            k32, _, _ = self.extractor.keys(images_u8[i:i + self.vit_chunk].contiguous(), want_f32=True)
            out.append(k32)
        return torch.cat(out, 0)

    @torch.no_grad()
    def get_features(self, originals: torch.Tensor, layout: str = "CHW"):
        """originals uint8 [N,3,H0,W0] -> (l keys fp32 [N,P,C], h keys fp32 [N,w*w,P,C], m keys | None) token-major;
        P = (S/patch)^2.  Resizes are Pillow-exact antialiased bilinear (transforms.Resize on the PIL image)."""
        N = originals.shape[0]
        H0, W0 = (originals.shape[-2:] if layout == "CHW" else originals.shape[1:3])
        dev = originals.device
        S, ws = self.S, self.ws
        jobs = torch.tensor([[n, 0, 0, W0, H0] for n in range(N)], dtype=torch.int32, device=dev)
        low = ops.roi_crop_resize(originals, jobs, (S, S), layout=layout)                     # [N,3,S,S]
        big = ops.roi_crop_resize(originals, jobs, (S * ws, S * ws), layout=layout)           # [N,3,3S,3S]
        wins = big.reshape(N, 3, ws, S, ws, S).permute(0, 2, 4, 1, 3, 5).reshape(N * ws * ws, 3, S, S)
        l = self._keys(low)
        h = self._keys(wins).reshape(N, ws * ws, l.shape[1], l.shape[2])
        m = None
        if self.require_m:
            # m-path: keys of the image at 54 * patch (756^2 for /14, 432^2 for /8), four overlapping 36^2 crops
            Sm = 54 * self.patch
            mid = ops.roi_crop_resize(originals, jobs, (Sm, Sm), layout=layout)
            km = self._keys(mid).reshape(N, 54, 54, -1)
            m = torch.stack([km[:, i * 18:i * 18 + 36, j * 18:j * 18 + 36] for i in range(2) for j in range(2)], 1)
            m = m.reshape(N, 4, 36 * 36, -1)
        return l, h, m

    @torch.no_grad()
    def _prepare(self, l, h, m):
        """loop_CORAL.py:206-245 on token-major keys."""
        N = l.shape[0]
        gp = int(round(l.shape[1] ** 0.5))
        g = self.g
        l_t, _ = ops.resize_tokens_bilinear(l, (gp, gp), (g, g))
        h_t, _ = ops.resize_tokens_bilinear(h.flatten(0, 1), (gp, gp), (g, g))
        h_t = h_t.reshape(N, self.ws * self.ws, g * g, -1)
        if self.require_m:
            _, m16 = ops.resize_tokens_bilinear(m.flatten(0, 1), (36, 36), (68, 68), want_f32=False, want_bf16=True)
            fg, _, _ = self.model.decoder.forward_tokens(m16, (68, 68), (68, 68), want_bg=False)
            preds = self.concate_preds(fg.reshape(N, 4, 1, 68, 68))
        else:
            _, l16 = ops.resize_tokens_bilinear(l, (gp, gp), (g, g), want_f32=False, want_bf16=True)
            preds, _, _ = self.model.decoder.forward_tokens(l16, (g, g), (g, g), want_bg=False)
        return l_t, h_t, preds

    @torch.no_grad()
    def refine(self, originals: torch.Tensor, layout: str = "CHW"):
        """-> (refined logits [N,1,.,.] list per image because centre-cropped images are padded to twice the size,
        crop flags)."""
        l, h, m = self.get_features(originals, layout)
        l_t, h_t, preds = self._prepare(l, h, m)
        crop = self._should_crop_center(preds)
        outputs, _, _ = self.refiner.forward_tokens(l_t, h_t, preds, self.g, per_image=True)
        results = [outputs[i:i + 1] for i in range(outputs.shape[0])]
        idx = torch.nonzero(crop).flatten().tolist()
        if idx:
            # low-confidence images: recompute everything on the centre half-size crop (loop_CORAL.py:276-311)
            sub = originals[idx]
            H0, W0 = (sub.shape[-2:] if layout == "CHW" else sub.shape[1:3])
            nh, nw = H0 // 2, W0 // 2
            top, left = (H0 - nh) // 2, (W0 - nw) // 2
            sub = (sub[..., top:top + nh, left:left + nw] if layout == "CHW" else sub[:, top:top + nh, left:left + nw])
            l2, h2, m2 = self.get_features(sub.contiguous(), layout)
            l2t, h2t, p2 = self._prepare(l2, h2, m2)
            out2, _, _ = self.refiner.forward_tokens(l2t, h2t, p2, self.g, per_image=True)
            out2 = self._center_pad(out2)
            for k, i in enumerate(idx):
                results[i] = out2[k:k + 1]
        return results, crop

    @torch.no_grad()
    def __call__(self, originals: torch.Tensor, label_sizes=None, layout: str = "CHW"):
        """-> list of uint8 masks, one per image, at `label_sizes[i]` (default: the original image size)."""
        results, _ = self.refine(originals, layout)
        H0, W0 = (originals.shape[-2:] if layout == "CHW" else originals.shape[1:3])
        return [self.process_preds(r, tuple(label_sizes[i]) if label_sizes is not None else (H0, W0))[0]
                for i, r in enumerate(results)]
