"""Look-Twice evaluation (reference: engine/runner/loop_UCOD_DPL.py:276-417, `ValLoop_Look_Twice`).

`LookTwiceEvaluator` is the batched device pipeline; `process_preds`, `expand_bbox`, `resize_bbox` and
`look_twice` keep the reference's per-image signatures so existing callers can switch over.
The reference needs the accelerate `runner` object only to reach `runner.model` and the config; here the
evaluator is constructed from those two things directly.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from ... import ops
from ...vit import VitKeyExtractor

DEFAULT_BOX = [129, 129, 259, 259]


def resize_bbox(bbox, original_width, original_height, new_width, new_height):
    """loop_UCOD_DPL.py:387-397 — python float64 + int() truncation (host side, a few boxes per image)."""
    x, y, w, h = bbox
    ws = new_width / original_width
    hs = new_height / original_height
    return [int(x * ws), int(y * hs), int(w * ws), int(h * hs)]


class LookTwiceEvaluator:
    """First look (ViT keys -> decoder@fs -> mask@S) -> boxes -> second look on the crops -> pasted mask."""

    def __init__(self, extractor: VitKeyExtractor, model, image_size, feature_size: int = 68,
                 look_twice_th: float = 0.15, expand_type: str = "dynamic", look_twice: bool = True):
        self.extractor = extractor
        self.model = model
        self.img_size = tuple(image_size)
        self.feature_size = int(feature_size)
        self.look_twice_th = float(look_twice_th)
        self.expand_type = expand_type
        self.enabled = look_twice
        self.patch = extractor.spec.patch

    # ---- reference-compatible pieces ----
    @torch.no_grad()
    def process_preds(self, preds: torch.Tensor, label_tensor=None):
        """preds [B,1,fs,fs] logits (CUDA).  B = 1: returns (preds_up [1,S,S] float, bboxes | None) exactly like the
        reference; B > 1: returns (preds_up [B,S,S] float, list of per-image bboxes | None)."""
        h, w = self.img_size
        mask = ops.upsample_bilinear(preds[:, 0], (h, w), binarize=True)
        boxes, nbox, status, _ = ops.lt_boxes(mask, self.look_twice_th, self.expand_type)
        nb = nbox.cpu().tolist()
        bx = boxes.cpu()
        out: List[Optional[list]] = []
        for b, n in enumerate(nb):
            if n == -2:
                raise ValueError("math domain error")  # expand_bbox: sqrt of a negative scale (reference raises too)
            out.append(None if n < 0 else bx[b, :n].tolist())
        up = mask.float()
        if preds.shape[0] == 1:
            return up, out[0]
        return up, out

    def expand_bbox(self, mask, bbox, img_width, img_height, expand_type="const", scale=1.3):
        """Host-side statement of loop_UCOD_DPL.py:399-417 for API parity (the device path computes the same in
        `ucod_lt_boxes`)."""
        x, y, w, h = bbox
        if expand_type == "dynamic":
            fr = float(mask[y:y + h, x:x + w].sum()) / (h * w)
            br = (h * y) / (mask.shape[-2] * mask.shape[-1])
            scale = math.sqrt(1 - br / fr + 1)
        new_w, new_h = w * scale, h * scale
        new_x = max(0, x - (new_w - w) / 2)
        if new_x + new_w > img_width:
            new_x = img_width - new_w
        new_y = max(0, y - (new_h - h) / 2)
        if new_y + new_h > img_height:
            new_y = img_height - new_h
        return [int(new_x), int(new_y), int(new_w), int(new_h)]

    resize_bbox = staticmethod(resize_bbox)

    @torch.no_grad()
    def first_look(self, images: torch.Tensor):
        _, k16, _ = self.extractor.keys(images, want_f32=False, want_bf16=True)
        gh, gw = images.shape[-2] // self.patch, images.shape[-1] // self.patch
        fs = self.feature_size
        fg, _, _ = self.model.decoder.forward_tokens(k16, (gh, gw), (fs, fs), want_bg=False)
        return fg

    @torch.no_grad()
    def look_twice_batch(self, originals: torch.Tensor, bboxes_per_image, masks_u8: torch.Tensor,
                         layout: str = "CHW", orig_sizes=None) -> torch.Tensor:
        """originals: uint8 RGB originals of the batch ([N,3,H0,W0] or [N,H0,W0,3]); bboxes_per_image: list (len N)
        of box lists or None; masks_u8 [N,S,S] {0,1}.  Returns new masks float [N,S,S] in [0,1]
        (loop_UCOD_DPL.py:326-352 for every image that has boxes; others keep their mask).
        orig_sizes (optional, [N,2] (h, w)): `originals` is a zero-padded canvas of differently sized images
        (`pack_padded`); boxes are mapped with each image's own size and crops past an image read 0 like PIL's."""
        ih, iw = self.img_size
        dev = masks_u8.device
        if layout == "CHW":
            H0, W0 = originals.shape[-2:]
        else:
            H0, W0 = originals.shape[1:3]
        sizes = None if orig_sizes is None else torch.as_tensor(orig_sizes).tolist()
        crop_jobs, paste_jobs = [], []
        for n, bxs in enumerate(bboxes_per_image):
            if bxs is None:
                continue
            if sizes is not None:
                H0, W0 = sizes[n]
            for rank, bb in enumerate(bxs):
                x, y, w, h = resize_bbox(bb, iw, ih, W0, H0)
                if w <= 0 or h <= 0 or bb[2] <= 0 or bb[3] <= 0:
                    raise ValueError("height and width must be > 0")  # PIL raises in the reference
                crop_jobs.append([n, x, y, w, h])
                paste_jobs.append([n, bb[0], bb[1], bb[2], bb[3], rank])
        canvas = ops.mask_scale_u8(masks_u8, 255)
        if crop_jobs:
            cj = torch.tensor(crop_jobs, dtype=torch.int32, device=dev)
            pj = torch.tensor(paste_jobs, dtype=torch.int32, device=dev)
            crops = ops.roi_crop_resize(originals, cj, (ih, iw), layout=layout)
            _, k16, _ = self.extractor.keys(crops, want_f32=False, want_bf16=True)
            g = (ih // self.patch, iw // self.patch)
            fg, _, _ = self.model.decoder.forward_tokens(k16, g, g, want_bg=False)  # raw 37^2 grid (:343-345)
            ops.paste_bicubic(fg[:, 0], pj, canvas)
        return canvas.float() / 255.0

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, originals: torch.Tensor | None = None, layout: str = "CHW",
                 orig_sizes=None):
        """images: network-size inputs [N,3,S,S] (uint8 raw or fp32 normalised); originals: the original-resolution
        uint8 images the crops are taken from (defaults to `images` when they are uint8; with `orig_sizes` a
        zero-padded canvas of ragged images).  Returns (final masks float [N,S,S] in [0,1], per-image bboxes)."""
        fg = self.first_look(images)
        up, bboxes = self.process_preds(fg)
        if images.shape[0] == 1:
            bboxes = [bboxes]
        if not self.enabled or all(b is None for b in bboxes):
            return up, bboxes
        if originals is None:
            if images.dtype != torch.uint8:
                raise ValueError("originals (uint8) are required when `images` are already normalised")
            originals = images
        mask_u8 = up.to(torch.uint8)
        return self.look_twice_batch(originals, bboxes, mask_u8, layout=layout, orig_sizes=orig_sizes), bboxes


class TrainLoop:
    """First-stage training schedule (reference: `TrainLoop`, engine/runner/loop_UCOD_DPL.py:35-255): per epoch
    [finetune switch] -> [discriminator epochs every `dis_intertrain` epochs until the finetune phase] -> decoder
    epoch -> [checkpoint] -> [Look-Twice validation, best MAE kept].  Same decision methods and counters.

    The reference pulls batches of cached features through a DataLoader; here the whole cache is resident in HBM
    (`keys` bf16 [N, gh*gw, dim] token-major, `pseudo_labels` [N,1,16,16]; 4040 training images @37x37 are 8.5 GB)
    and an epoch is a seeded device permutation cut into batches, each rank taking every world-th batch slot
    (gradients are all-reduced inside `FirstStageTrainer.process_batch`).
    """

    def __init__(self, config, trainer, dis_trainer, keys: torch.Tensor, pseudo_labels: torch.Tensor, grid_in,
                 validate=None, save_checkpoint=None, logger=None, seed: int = 0, rank: int = 0, world_size: int = 1):
        self.cfg = config
        self.trainer, self.dis_trainer = trainer, dis_trainer
        self.keys, self.pl, self.grid_in = keys, pseudo_labels, tuple(grid_in)
        self.validate, self.save_checkpoint, self.logger = validate, save_checkpoint, logger
        self.rank, self.world = rank, world_size
        tc = config.train_cfg
        self._start_epoch = tc.start_epoch
        self._max_epoch = tc.max_epoch
        self._cur_epoch = 0
        self._start_finetune = tc.start_finetune
        self.finetune = False
        self.global_step = 0
        self.batch_size = int(config.dataset_cfg.trainloader_cfg.batch_size)
        self.shuffle = bool(config.dataset_cfg.trainloader_cfg.shuffle)
        self.enable_val = config.val_cfg.enable_val
        self.val_interval = config.val_cfg.val_interval
        self.dis_intertrain = tc.dis_intertrain
        sv = config.val_cfg.start_val
        self.val_start = self._max_epoch + sv if sv < 0 else sv
        ss = tc.save_cfg.start_save
        self.save_start = self._max_epoch + ss if ss < 0 else ss
        self.save_interval = tc.save_cfg.save_interval
        self.log_interval = config.log_cfg.log_interval
        self.best_mae = 1000.0
        self.best_result = None
        self.losses: List[float] = []
        self._gen = torch.Generator(device="cpu").manual_seed(seed)

    # ---- schedule decisions (loop_UCOD_DPL.py:193-215) ----
    def decide_to_train_dis(self) -> bool:
        if self.cfg.train_cfg.merge_method != "dis":
            return False
        return self._cur_epoch % self.dis_intertrain == 0 and not self.finetune

    def decide_to_finetune(self) -> bool:
        if self._cur_epoch == self._max_epoch + self._start_finetune:
            self.finetune = True
            return True
        return False

    def decide_to_save(self) -> bool:
        return self._cur_epoch >= self.save_start and self._cur_epoch % self.save_interval == 0

    def decide_to_val(self) -> bool:
        return bool(self.enable_val and self._cur_epoch >= self.val_start and self._cur_epoch % self.val_interval == 0)

    # ---- data ----
    def _batches(self):
        """this rank's batches of one epoch: index tensors on the device."""
        n = self.keys.shape[0]
        order = torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)
        per_step = self.batch_size * self.world
        for s in range(0, n, per_step):
            idx = order[s + self.rank * self.batch_size: s + (self.rank + 1) * self.batch_size]
            if idx.numel() == 0:       # ragged tail: ranks without data repeat the head of the permutation, so that
                idx = order[: self.batch_size]  # every rank joins every gradient all-reduce
            yield idx.to(self.keys.device)

    def _log(self, msg: str) -> None:
        if self.logger is not None:
            self.logger.info(msg)

    # ---- epochs ----
    def Discriminator_epoch(self) -> None:
        for idx in self._batches():
            loss = self.dis_trainer.epoch_step(self.trainer.model, self.keys.index_select(0, idx), self.grid_in,
                                               self.pl.index_select(0, idx), feature_size=self.trainer.fs)
            if self._cur_epoch % self.log_interval == 0:
                self._log("dis:loss:{:.4f}".format(float(loss)))

    def Discriminator_train(self) -> None:
        for _ in range(self.cfg.train_cfg.dis_epoch):
            self.Discriminator_epoch()

    def run_epoch(self) -> None:
        self.trainer.cur_epoch = self._cur_epoch
        for idx in self._batches():
            loss = self.trainer.process_batch(self.keys.index_select(0, idx), self.grid_in, self.pl.index_select(0, idx))
            if self._cur_epoch % self.log_interval == 0:
                lv = float(loss)
                self.losses.append(lv)
                self._log(f"iter{self.global_step}:loss:{lv:.4f}")
            self.global_step += 1

    def _update_best_result(self, result) -> None:
        if result["MAE"] < self.best_mae:
            self.best_mae = result["MAE"]
            self.best_result = result
            self._log("best result: {}".format({k: [round(float(v), 4)] for k, v in result.items()}))

    def run(self):
        while self._cur_epoch < self._max_epoch:
            if self.decide_to_finetune():
                self.trainer.start_finetune_phase()
                if self.dis_trainer is not None:
                    self.dis_trainer.reset_optimizer()
                self.global_step = 0
            if self.decide_to_train_dis():
                self.Discriminator_train()
            self.run_epoch()
            self._cur_epoch += 1
            if self.decide_to_save() and self.save_checkpoint is not None:
                self.save_checkpoint(self._cur_epoch)
            if self.decide_to_val() and self.validate is not None:
                self._update_best_result(self.validate())
        return self.best_result
