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


class LookTwiceResult:
    """What one Look-Twice pass leaves on the device: nothing here has synchronised with the host.

    final     float [N,S,S] in [0,1] — the reference's `preds_up` after the second look (loop_UCOD_DPL.py:352)
    first     uint8 [N,S,S] {0,1} — the first-look mask (`preds_up` of process_preds)
    boxes / nbox   int32 [N,128,4] / [N] as written by `ucod_lt_boxes` (nbox -1 = None, -2 = ValueError)
    counts    int32 [4] (second looks kept, status bits, second looks requested, 0) or None when Look-Twice is off
    `check()` is the one place that reads anything back (a few dozen bytes): it raises what the reference would have
    raised for this batch, and — if the batch needed more second-look chunks than were enqueued ahead of time, or more
    second looks than the device job table holds — completes the batch and refreshes `final` (see
    `LookTwiceEvaluator.look_twice_device`).
    `bboxes` converts the box table to the reference's per-image lists (also a host read)."""

    def __init__(self, final, first, boxes, nbox, counts, err, canvas=None, resume=None, redo=None):
        self.final, self.first, self.boxes, self.nbox, self.counts, self.err = final, first, boxes, nbox, counts, err
        self.canvas, self._resume, self._redo = canvas, resume, redo
        self._lists = None

    def check(self) -> None:
        nb = self.nbox.cpu()
        if bool((nb == -3).any()) and self._redo is not None:
            # a mask exceeded the shared-memory labeller's run capacity (never the case for masks up-sampled from the
            # decoder's logit grid): redo the batch with the global-memory labeller
            other = self._redo()
            other.check()
            for name in ("final", "first", "boxes", "nbox", "counts", "err", "canvas"):
                setattr(self, name, getattr(other, name))
            self._resume = self._redo = None
            return
        if bool((nb == -2).any()):
            raise ValueError("math domain error")  # expand_bbox: sqrt of a negative scale (reference raises too)
        if self.counts is not None:
            kept, status, wanted, _ = self.counts.cpu().tolist()
            if status & 4:
                raise ValueError("height and width must be > 0")  # PIL raises in the reference
            if self._resume is not None:
                self._resume(self, kept, wanted)
                self._resume = None
        if self.err is not None and int(self.err.item()):
            raise ops.UcodError("Look-Twice resampling: a crop or box is outside the supported scale range")

    @property
    def bboxes(self):
        if self._lists is None:
            nb = self.nbox.cpu().tolist()
            bx = self.boxes.cpu()
            self._lists = [None if n < 0 else bx[b, :n].tolist() for b, n in enumerate(nb)]
        return self._lists


class LookTwiceEvaluator:
    """First look (ViT keys -> decoder@fs -> mask@S) -> boxes -> second look on the crops -> pasted mask.

    The whole flow stays on the device: box lists become crop / paste job tables in a kernel (`ucod_lt_build_jobs`),
    the second ViT + decoder pass runs on fixed-size chunks of that table whose valid length is read by the kernels
    themselves (`*_dyn` entry points), and pastes resolve their order per pixel.  The host only enqueues."""

    def __init__(self, extractor: VitKeyExtractor, model, image_size, feature_size: int = 68,
                 look_twice_th: float = 0.15, expand_type: str = "dynamic", look_twice: bool = True,
                 max_looks_per_image: int = 4):
        self.extractor = extractor
        self.model = model
        self.img_size = tuple(image_size)
        self.feature_size = int(feature_size)
        self.look_twice_th = float(look_twice_th)
        self.expand_type = expand_type
        self.enabled = look_twice
        self.max_looks_per_image = int(max_looks_per_image)
        self.patch = extractor.spec.patch
        # second-look chunks enqueued ahead of time (see look_twice_device); adapts to what recent batches needed
        self._recent_chunks: list = []

    # ---- reference-compatible pieces ----
    @torch.no_grad()
    def process_preds(self, preds: torch.Tensor, label_tensor=None):
        """preds [B,1,fs,fs] logits (CUDA).  B = 1: returns (preds_up [1,S,S] float, bboxes | None) exactly like the
        reference; B > 1: returns (preds_up [B,S,S] float, list of per-image bboxes | None).  Host lists mean a
        device->host read; the batched pipeline (`look_twice_device`) never calls this."""
        h, w = self.img_size
        mask = ops.upsample_bilinear(preds[:, 0], (h, w), binarize=True)
        boxes, nbox, status, _ = ops.lt_boxes(mask, self.look_twice_th, self.expand_type)
        if bool((nbox == -3).any()):  # labeller capacity (see ops.lt_boxes)
            boxes, nbox, status, _ = ops.lt_boxes(mask, self.look_twice_th, self.expand_type, algorithm="global")
        res = LookTwiceResult(None, mask, boxes, nbox, None, None)
        res.check()
        up = mask.float()
        if preds.shape[0] == 1:
            return up, res.bboxes[0]
        return up, res.bboxes

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
        """Second look for HOST box lists (the reference's `look_twice` signature, batched): originals uint8 RGB
        ([N,3,H0,W0] or [N,H0,W0,3]); bboxes_per_image: list (len N) of box lists or None; masks_u8 [N,S,S] {0,1}.
        Returns new masks float [N,S,S] in [0,1] (loop_UCOD_DPL.py:326-352 for every image that has boxes; others keep
        their mask).  orig_sizes (optional, [N,2] (h, w)): `originals` is a zero-padded canvas of differently sized
        images (`pack_padded`); boxes are mapped with each image's own size and crops past an image read 0 like PIL's.
        The device-resident pipeline is `look_twice_device`; this entry point exists for callers that hold lists."""
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
        return ops.to_tensor_normalize(canvas[:, None])[:, 0]

    @torch.no_grad()
    def look_twice_device(self, images: torch.Tensor, originals: torch.Tensor | None = None, layout: str = "CHW",
                          orig_sizes=None, first_logits: torch.Tensor | None = None,
                          ccl_algorithm: str = "auto") -> LookTwiceResult:
        """The whole of loop_UCOD_DPL.py:297-313 for a batch, enqueued without a single host synchronisation.
        images: network-size inputs [N,3,S,S] (uint8 raw or fp32 normalised); originals: the original-resolution
        uint8 images the crops are taken from (defaults to `images` when they are uint8; with `orig_sizes` [N,2] a
        zero-padded canvas of ragged images); first_logits: optional [N,1,fs,fs] logits that replace the first-look
        prediction as `process_preds`' input (the first look still runs; used to plant objects in benchmarks).

        The second look runs on chunks of N crops whose valid length the kernels read on the device.  How many chunks
        are enqueued ahead of time follows what the last batches needed (an empty chunk still costs ~60 early-exit
        launches); a batch that needs more is completed by `LookTwiceResult.check()`, which every consumer calls when
        it reads the masks back — results never depend on the guess, only the launch count does."""
        fg = self.first_look(images)
        if first_logits is not None:
            fg = first_logits
        ih, iw = self.img_size
        mask = ops.upsample_bilinear(fg[:, 0], (ih, iw), binarize=True)
        boxes, nbox, _, _ = ops.lt_boxes(mask, self.look_twice_th, self.expand_type, algorithm=ccl_algorithm)
        canvas = ops.mask_scale_u8(mask, 255)
        redo = None
        if ccl_algorithm != "global":
            redo = lambda: self.look_twice_device(images, originals, layout, orig_sizes, first_logits, "global")  # noqa: E731
        if not self.enabled:
            return LookTwiceResult(ops.to_tensor_normalize(canvas[:, None])[:, 0], mask, boxes, nbox, None, None,
                                   redo=redo)
        if originals is None:
            if images.dtype != torch.uint8:
                raise ValueError("originals (uint8) are required when `images` are already normalised")
            originals = images
        N = images.shape[0]
        H0, W0 = originals.shape[-2:] if layout == "CHW" else originals.shape[1:3]
        chunk = max(N, 16)
        capacity = chunk * self.max_looks_per_image
        sizes = None if orig_sizes is None else torch.as_tensor(orig_sizes)
        crop, paste, counts, chunks = ops.lt_build_jobs(boxes, nbox, (ih, iw), (H0, W0), sizes, capacity=capacity,
                                                        chunk=chunk)
        err = torch.zeros(1, device=mask.device, dtype=torch.int32)
        g = (ih // self.patch, iw // self.patch)
        out_cap = int(math.ceil(1.5 * max(ih, iw)))  # expand_bbox grows a box by at most sqrt(2)

        def run_chunk(c):
            n_c = chunks[c:c + 1]
            crops = ops.roi_crop_resize_dyn(originals, crop[c * chunk:(c + 1) * chunk], n_c, (ih, iw), layout=layout,
                                            err=err)
            _, k16, _ = self.extractor.keys(crops, want_f32=False, want_bf16=True, count_dev=n_c)
            fg2, _, _ = self.model.decoder.forward_tokens(k16, g, g, want_bg=False, count_dev=n_c)  # :343-345
            ops.paste_bicubic_dyn(fg2[:, 0], paste, c * chunk, n_c, counts[0:1], canvas, out_cap, err=err)

        ahead = min(self.max_looks_per_image, max(self._recent_chunks)) if self._recent_chunks else self.max_looks_per_image
        for c in range(ahead):
            run_chunk(c)
        final = ops.to_tensor_normalize(canvas[:, None])[:, 0]  # / 255 (:352)

        def resume(res, kept, wanted):  # called from check(): the host knows the counts now
            need = (kept + chunk - 1) // chunk
            self._recent_chunks = (self._recent_chunks + [need])[-8:]
            if wanted > kept:
                # more second looks than the job table holds (>= 16 * max_looks_per_image): redo the second look from
                # host box lists, which have no capacity (the reference's own loop shape; rare by construction)
                res.final.copy_(self.look_twice_batch(originals, res.bboxes, mask, layout=layout,
                                                      orig_sizes=orig_sizes))
            elif need > ahead:
                for c in range(ahead, need):
                    run_chunk(c)
                res.final.copy_(ops.to_tensor_normalize(canvas[:, None])[:, 0])

        return LookTwiceResult(final, mask, boxes, nbox, counts, err, canvas=canvas, resume=resume, redo=redo)

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, originals: torch.Tensor | None = None, layout: str = "CHW",
                 orig_sizes=None):
        """Returns (final masks float [N,S,S] in [0,1], per-image bboxes) like the reference's loop body; the box
        lists and the error checks are read back once, after everything has been enqueued."""
        res = self.look_twice_device(images, originals, layout, orig_sizes)
        res.check()
        return res.final, res.bboxes


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
        if self.world > 1 and n % per_step:
            # accelerate's BatchSamplerShard(even_batches=True): the last global batch is completed by wrapping around
            # to the start of the permutation, so every rank contributes a full batch to every gradient all-reduce and
            # no sample is weighted more than another within a step
            pad = per_step - n % per_step
            order = torch.cat([order, order[:pad]] if pad <= n else [order] + [order] * (pad // n) + [order[: pad % n]])
        for s in range(0, order.numel(), per_step):
            idx = order[s + self.rank * self.batch_size: s + (self.rank + 1) * self.batch_size]
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

    def _sync_discriminator_buffers(self) -> None:
        """DDP broadcasts module buffers from rank 0 at every forward (broadcast_buffers=True); the BatchNorm running
        statistics of the discriminator are per-rank here, so they are aligned once per discriminator epoch."""
        if self.world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            for buf in self.dis_trainer.D.buffers():
                torch.distributed.broadcast(buf, src=0)

    def Discriminator_train(self) -> None:
        for _ in range(self.cfg.train_cfg.dis_epoch):
            self.Discriminator_epoch()
            self._sync_discriminator_buffers()

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
