"""First-stage training step on the GPU (reference: engine/runner/loop_UCOD_DPL.py:36-272 `TrainLoop`).

* `decoder_forward_autograd` — `RevDecoder.forward` in training mode as a `torch.autograd.Function` whose backward
  is the hand-written CUDA backward (`ucod_decoder_bwd`), so `loss.backward()` on (fg, bg, ortho) works with any
  optimiser exactly like the reference module.
* `FirstStageTrainer`       — the fused step: teacher (EMA) forward, student forward, APM merge, BCE + ortho loss
  and backward in one kernel sequence, decoder-gradient all-reduce over NCCL (flat 98 690-float buffer), fused
  AdamW + EMA update, StepLR stepped per iteration.  Mirrors `_process_batch` / `update_ema_decoder`.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from ._lib import c_float, ptr, stream_ptr
from .models.discriminator import merge_pseudo_label

_u64 = ctypes.c_uint64


def _decoder_pointers(dec):
    return (ptr(dec._w_dec_bf16()), ptr(dec.decoupling.bias.detach()), ptr(dec.learnable_embedding.detach()),
            ptr(dec.conv_out_fg.weight.detach()), ptr(dec.conv_out_fg.bias.detach()),
            ptr(dec.conv_out_bg.weight.detach()), ptr(dec.conv_out_bg.bias.detach()))


def _forward_with_workspace(dec, tokens, grid_in, grid_out):
    """Student forward that keeps its workspace (d_in, sumsq, fhat, Grams) for the backward."""
    B, P, dim = tokens.shape
    (gh, gw), (oh, ow) = grid_in, grid_out
    dev = tokens.device
    lib = _lib.load()
    lib.ucod_decoder_workspace_bytes.restype = _u64
    need = lib.ucod_decoder_workspace_bytes(B, gh, gw, oh, ow, 1)
    ws = torch.empty(int(need) + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    fg = torch.empty(B, 1, oh, ow, device=dev)
    bg = torch.empty(B, 1, oh, ow, device=dev)
    ortho = torch.empty((), device=dev)
    with torch.cuda.device(dev):
        _lib.call("ucod_decoder_fwd", ptr(tokens), B, dim, gh, gw, oh, ow, *_decoder_pointers(dec), ptr(fg), ptr(bg),
                  ptr(ortho), ctypes.c_void_p(ws.data_ptr() + off), _u64(ws.numel() - off), stream_ptr(dev))
    return fg, bg, ortho, (ws, off)


def _backward(dec, tokens, grid_in, grid_out, fg, bg, fwd_ws, grads, *, target=None, dfg=None, dbg=None, dortho=None):
    """grads: dict name -> fp32 tensor views (w_dec [128,dim], b_dec, w_fg, b_fg, w_bg, b_bg). Returns loss2 | None."""
    B, P, dim = tokens.shape
    (gh, gw), (oh, ow) = grid_in, grid_out
    dev = tokens.device
    lib = _lib.load()
    lib.ucod_decoder_bwd_workspace_bytes.restype = _u64
    bws = torch.empty(int(lib.ucod_decoder_bwd_workspace_bytes(B, gh, gw, oh, ow)) + 1024, dtype=torch.uint8, device=dev)
    boff = (-bws.data_ptr()) % 1024
    ws, off = fwd_ws
    loss2 = torch.empty(2, device=dev)
    with torch.cuda.device(dev):
        _lib.call("ucod_decoder_bwd", ptr(tokens), B, dim, gh, gw, oh, ow, *_decoder_pointers(dec), ptr(fg), ptr(bg),
                  ptr(target), ptr(dfg), ptr(dbg), ptr(dortho), ctypes.c_void_p(ws.data_ptr() + off),
                  _u64(ws.numel() - off), ptr(grads["w_dec"]), ptr(grads["b_dec"]), ptr(grads["w_fg"]),
                  ptr(grads["b_fg"]), ptr(grads["w_bg"]), ptr(grads["b_bg"]), ptr(loss2),
                  ctypes.c_void_p(bws.data_ptr() + boff), _u64(bws.numel() - boff), stream_ptr(dev))
    return loss2


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, tokens, grid, w, b, emb, wfg, bfg, wbg, bbg):
        fg, bg, ortho, ws = _forward_with_workspace(dec, tokens, grid, grid)
        ctx.dec, ctx.tokens, ctx.grid, ctx.ws = dec, tokens, grid, ws
        ctx.save_for_backward(fg, bg)
        return fg, bg, ortho

    @staticmethod
    def backward(ctx, dfg, dbg, dortho):
        dec, tokens = ctx.dec, ctx.tokens
        fg, bg = ctx.saved_tensors
        dev = tokens.device
        dim = tokens.shape[-1]
        g = {"w_dec": torch.empty(128, dim, device=dev), "b_dec": torch.empty(128, device=dev),
             "w_fg": torch.empty(64, device=dev), "b_fg": torch.empty(1, device=dev),
             "w_bg": torch.empty(64, device=dev), "b_bg": torch.empty(1, device=dev)}
        zeros = lambda t: torch.zeros_like(t) if t is None else t.contiguous().float()  # noqa: E731
        dfg = torch.zeros_like(fg) if dfg is None else dfg.contiguous().float()
        dbg = torch.zeros_like(bg) if dbg is None else dbg.contiguous().float()
        dortho = torch.zeros((), device=dev) if dortho is None else dortho.contiguous().float()
        _backward(dec, tokens, ctx.grid, ctx.grid, fg, bg, ctx.ws, g, dfg=dfg, dbg=dbg, dortho=dortho)
        return (None, None, None, g["w_dec"].reshape(128, dim, 1, 1), g["b_dec"],
                torch.zeros_like(dec.learnable_embedding), g["w_fg"].reshape(1, 64, 1, 1), g["b_fg"],
                g["w_bg"].reshape(1, 64, 1, 1), g["b_bg"])


def decoder_forward_autograd(dec, tokens_bf16: torch.Tensor, grid):
    """(fg, bg, ortho) with autograd support for the student decoder parameters (features need no gradient)."""
    return _DecoderFn.apply(dec, tokens_bf16, tuple(grid), dec.decoupling.weight, dec.decoupling.bias,
                            dec.learnable_embedding, dec.conv_out_fg.weight, dec.conv_out_fg.bias,
                            dec.conv_out_bg.weight, dec.conv_out_bg.bias)


class FirstStageTrainer:
    """Fused first-stage training (config 5).  `model` is a `baseline`; `discriminator` a `Discriminator` (frozen
    during decoder epochs, BatchNorm in train mode like the reference)."""

    def __init__(self, model, discriminator, *, lr0: float = 2e-4, step_lr_size: int = 25, step_lr_gamma: float = 0.95,
                 feature_size: int = 68, ema_weight: float = 0.99, max_epoch: int = 25, start_finetune: int = -5,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, process_group=None,
                 use_graph: bool = False):
        self.model, self.discriminator = model, discriminator
        # CUDA-graph replay of the forward / APM / backward kernel sequence (~45 launches and 9 memsets per step, each
        # a few microseconds of GPU time: eagerly the step is bound by the host enqueueing them).  The graph is
        # captured after two eager steps and re-captured when the batch geometry, the epoch (the APM schedule term is
        # a launch argument) or the finetune flag change; the gradient all-reduce and the AdamW / EMA kernel, whose
        # scalars change every step, stay outside it.
        self.use_graph = bool(use_graph)
        self._graph = None
        self._graph_key = None
        self._eager_left = 0
        self.fs = int(feature_size)
        self.lr0, self.step_size, self.gamma = lr0, step_lr_size, step_lr_gamma
        self.ema_weight, self.max_epoch, self.start_finetune = ema_weight, max_epoch, start_finetune
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        self.pg = process_group
        self.global_step = 0      # the reference's counter (bumped twice per batch: loop_UCOD_DPL.py:143,182)
        self.opt_steps = 0        # optimiser / scheduler steps
        self.cur_epoch = 0
        self.finetune = False
        dec, ema = model.decoder, model.decoder_ema
        dev = dec.decoupling.weight.device
        if dev.type != "cuda":
            raise _lib.UcodError("FirstStageTrainer needs the model on a CUDA device; there is no CPU fallback")
        # flat parameter / EMA / gradient / moment buffers in nn.Module.parameters() order; the modules' parameters
        # become views of the flat buffers, so state_dict() / checkpoints are unaffected.
        # Every slot is padded to a multiple of 4 floats so that the kernels' 128-bit loads stay aligned.
        self.names = [n for n, _ in dec.named_parameters()]
        sizes = [p.numel() for p in dec.parameters()]
        slots = [(sz + 3) // 4 * 4 for sz in sizes]
        self.n = sum(slots)
        self.flat_p = torch.zeros(self.n, device=dev)
        self.flat_ema = torch.zeros(self.n, device=dev)
        self.flat_g = torch.zeros(self.n, device=dev)
        self.flat_m = torch.zeros(self.n, device=dev)
        self.flat_v = torch.zeros(self.n, device=dev)
        self.views = {}
        o = 0
        for (name, p), (_, pe), sz, slot in zip(dec.named_parameters(), ema.named_parameters(), sizes, slots):
            self.flat_p[o:o + sz].copy_(p.detach().flatten())
            self.flat_ema[o:o + sz].copy_(pe.detach().flatten())
            p.data = self.flat_p[o:o + sz].view_as(p)
            pe.data = self.flat_ema[o:o + sz].view_as(pe)
            self.views[name] = self.flat_g[o:o + sz]
            o += slot
        dim = dec.decoupling.weight.shape[1]
        self.grads = {"w_dec": self.views["decoupling.weight"].view(128, dim), "b_dec": self.views["decoupling.bias"],
                      "w_fg": self.views["conv_out_fg.weight"], "b_fg": self.views["conv_out_fg.bias"],
                      "w_bg": self.views["conv_out_bg.weight"], "b_bg": self.views["conv_out_bg.bias"]}

    @property
    def lr(self) -> float:
        return self.lr0 * self.gamma ** (self.opt_steps // self.step_size)

    def start_finetune_phase(self):
        """`Runner.start_finetune` + `TrainLoop.run` (runner.py:378-379, loop_UCOD_DPL.py:101-103): the optimiser and
        the StepLR schedule are rebuilt (fresh AdamW moments, lr back to lr0) and global_step restarts, which also
        restarts the EMA warm-up `1 - 1/(step+1)`."""
        self.finetune = True
        self.global_step = 0
        self.opt_steps = 0
        self.flat_m.zero_()
        self.flat_v.zero_()

    @torch.no_grad()
    def _forward_backward(self, key_tokens_bf16: torch.Tensor, grid_in, pseudo_labels: torch.Tensor):
        """teacher fwd, student fwd, APM merge, BCE + ortho loss, decoder backward into `self.flat_g`.
        Returns (loss, dict of intermediates); only enqueues device work (graph-capturable)."""
        dec, ema = self.model.decoder, self.model.decoder_ema
        fs = (self.fs, self.fs)
        pl = ops.upsample_bilinear(pseudo_labels.float(), fs)                    # F.interpolate(pl, 68x68)
        # both decoders' weights are rewritten in place by the fused AdamW / EMA kernel (raw pointers: torch's version
        # counters do not move), so their cached bf16 copies of the 1x1 conv weight are rebuilt every step
        dec._packed = ema._packed = None
        teacher, _, _ = ema.forward_tokens(key_tokens_bf16, grid_in, fs, want_bg=False)
        fg, bg, ortho, ws = _forward_with_workspace(dec, key_tokens_bf16, grid_in, fs)
        merged, dis_loss = merge_pseudo_label(self.discriminator, pl, teacher, fg, None, cur_epoch=self.cur_epoch,
                                              max_epoch=self.max_epoch, start_finetune=self.start_finetune)
        loss2 = _backward(dec, key_tokens_bf16, grid_in, fs, fg, bg, ws, self.grads, target=merged)
        # (the gradient slot of learnable_embedding is never written: it stays at the zeros it was created with)
        loss = torch.empty((), device=loss2.device)
        with torch.cuda.device(loss2.device):  # loss = bce_fg + bce_bg + ortho [- dis_loss], one launch
            _lib.call("ucod_train_loss", ptr(loss2), ptr(ortho), ptr(None if self.finetune else dis_loss), ptr(loss),
                      stream_ptr(loss2.device))
        return loss, {"merged": merged, "dis_loss": dis_loss, "ortho": ortho, "bce": loss2}

    @torch.no_grad()
    def _optimizer_step(self) -> None:
        world = 1
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(self.pg)
            if world > 1:
                torch.distributed.all_reduce(self.flat_g, group=self.pg)         # one 395 KB NCCL all-reduce
        self.opt_steps += 1
        lr = self.lr0 * self.gamma ** ((self.opt_steps - 1) // self.step_size)
        alpha = min(1 - 1 / (self.global_step + 1), self.ema_weight)
        dev = self.flat_p.device
        with torch.cuda.device(dev):
            _lib.call("ucod_adamw_ema_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), ptr(self.flat_v),
                      ptr(self.flat_ema), _u64(self.n), c_float(lr), c_float(self.betas[0]), c_float(self.betas[1]),
                      c_float(self.eps), c_float(self.wd), int(self.opt_steps), c_float(1.0 / world), c_float(alpha),
                      stream_ptr(dev))
        self.global_step += 2

    @torch.no_grad()
    def process_batch(self, key_tokens_bf16: torch.Tensor, grid_in, pseudo_labels: torch.Tensor):
        """key_tokens_bf16 [B, gh*gw, dim] (cached backbone keys), pseudo_labels [B,1,16,16] {0,1}.
        Returns the loss tensor (device scalar), like `_process_batch`."""
        if not self.use_graph:
            loss, self.last = self._forward_backward(key_tokens_bf16, grid_in, pseudo_labels)
            self._optimizer_step()
            return loss
        key = (tuple(key_tokens_bf16.shape), tuple(pseudo_labels.shape), tuple(grid_in), self.cur_epoch, self.finetune)
        if key != self._graph_key:
            self._graph, self._graph_key, self._eager_left = None, key, 2
        if self._graph is None and self._eager_left > 0:       # real steps, also warm every kernel / allocation up
            self._eager_left -= 1
            loss, self.last = self._forward_backward(key_tokens_bf16, grid_in, pseudo_labels)
            self._optimizer_step()
            return loss
        if self._graph is None:
            self._s_tok = key_tokens_bf16.clone()
            self._s_pl = pseudo_labels.clone()
            torch.cuda.synchronize(self.flat_p.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._s_loss, self._s_last = self._forward_backward(self._s_tok, grid_in, self._s_pl)
            self._graph = g
        # a caller that assembles its batch directly in the graph's input buffers (`graph_inputs()`; e.g.
        # `torch.index_select(store, 0, idx, out=tok_buf)` from an HBM-resident training set) skips these copies
        if key_tokens_bf16.data_ptr() != self._s_tok.data_ptr():
            self._s_tok.copy_(key_tokens_bf16)
        if pseudo_labels.data_ptr() != self._s_pl.data_ptr():
            self._s_pl.copy_(pseudo_labels)
        self._graph.replay()
        self.last = self._s_last
        self._optimizer_step()
        return self._s_loss


    def graph_inputs(self):
        """(key-token buffer, pseudo-label buffer) the captured graph reads, or None before the capture."""
        return None if self._graph is None else (self._s_tok, self._s_pl)


class _DiscPtrs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "conv1", "bn1_w", "bn1_b", "bn1_mean", "bn1_var", "conv2", "bn2_w", "bn2_b", "bn2_mean", "bn2_var",
        "conv3", "bn3_w", "bn3_b", "bn3_mean", "bn3_var", "lin_w", "lin_b")]


class _DiscGradPtrs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("conv1", "bn1_w", "bn1_b", "conv2", "bn2_w", "bn2_b", "conv3", "bn3_w",
                                               "bn3_b", "lin_w", "lin_b")]


class DiscriminatorTrainer:
    """`TrainLoop.Discriminator_epoch` iterations (engine/runner/loop_UCOD_DPL.py:230-255): BCE on
    [student masks -> 0, pseudo labels -> 1], AdamW(dis_lr0) + StepLR per iteration.  Forward and backward run in
    csrc/discriminator.cu; parameters live in one flat buffer (the module's parameters are views of it)."""

    _GRAD_FIELDS = ["conv1", "bn1_w", "bn1_b", "conv2", "bn2_w", "bn2_b", "conv3", "bn3_w", "bn3_b", "lin_w", "lin_b"]

    def __init__(self, discriminator, *, lr0: float = 1e-3, step_lr_size: int = 25, step_lr_gamma: float = 0.95,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, process_group=None):
        self.D = discriminator
        self.lr0, self.step_size, self.gamma = lr0, step_lr_size, step_lr_gamma
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        self.pg = process_group
        self.opt_steps = 0
        params = list(discriminator.parameters())     # maskConv.{conv,bn.w,bn.b}, convs.0.*, convs.1.*, linear.{w,b}
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.UcodError("DiscriminatorTrainer needs the discriminator on a CUDA device")
        sizes = [p.numel() for p in params]
        slots = [(s + 3) // 4 * 4 for s in sizes]
        self.n = sum(slots)
        self.flat_p = torch.zeros(self.n, device=dev)
        self.flat_g = torch.zeros(self.n, device=dev)
        self.flat_m = torch.zeros(self.n, device=dev)
        self.flat_v = torch.zeros(self.n, device=dev)
        self.grad_views = []
        o = 0
        for p, sz, slot in zip(params, sizes, slots):
            self.flat_p[o:o + sz].copy_(p.detach().flatten())
            p.data = self.flat_p[o:o + sz].view_as(p)
            self.grad_views.append(self.flat_g[o:o + sz])
            o += slot
        self.loss = torch.zeros((), device=dev)

    def reset_optimizer(self):
        """`_build_optimizer` again (runner.py:276-308, called by `start_finetune`): fresh moments and schedule."""
        self.opt_steps = 0
        self.flat_m.zero_()
        self.flat_v.zero_()

    def _forward(self, mask):
        D = self.D
        B, _, fs, _ = mask.shape
        dev = mask.device
        lib = _lib.load()
        lib.ucod_discriminator_workspace_bytes.restype = _u64
        ws = torch.empty(int(lib.ucod_discriminator_workspace_bytes(B, fs)) + 1024, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 1024
        t = D._tensors()
        w = _DiscPtrs(*[t[n].data_ptr() for n, _ in _DiscPtrs._fields_])
        prob = torch.empty(B, 1, device=dev)
        with torch.cuda.device(dev):
            _lib.call("ucod_discriminator_fwd", ptr(mask), B, fs, ctypes.byref(w), 1, 1, ptr(prob),
                      ctypes.c_void_p(ws.data_ptr() + off), _u64(ws.numel() - off), stream_ptr(dev))
        for blk in (D.maskConv.layers, D.convs[0].layers, D.convs[1].layers):
            blk[1].num_batches_tracked += 1
        return prob, w, (ws, off)

    def _backward(self, mask, prob, w, fwd_ws, label: float, n_total: int):
        B, _, fs, _ = mask.shape
        dev = mask.device
        lib = _lib.load()
        lib.ucod_discriminator_bwd_workspace_bytes.restype = _u64
        bws = torch.empty(int(lib.ucod_discriminator_bwd_workspace_bytes(B, fs)) + 1024, dtype=torch.uint8, device=dev)
        boff = (-bws.data_ptr()) % 1024
        g = _DiscGradPtrs(*[v.data_ptr() for v in self.grad_views])
        ws, off = fwd_ws
        with torch.cuda.device(dev):
            _lib.call("ucod_discriminator_bwd", ptr(mask), B, fs, ctypes.byref(w), ptr(prob), c_float(label),
                      int(n_total), ctypes.byref(g), ptr(self.loss), ctypes.c_void_p(ws.data_ptr() + off),
                      ctypes.c_void_p(bws.data_ptr() + boff), _u64(bws.numel() - boff), stream_ptr(dev))

    @torch.no_grad()
    def step(self, pseudo_masks: torch.Tensor, student_masks: torch.Tensor):
        """pseudo_masks, student_masks: fp32 {0,1} [B,1,fs,fs] (already binarised like loop_UCOD_DPL.py:239-240).
        Returns the BCE loss (device scalar)."""
        pm, sm = pseudo_masks.float().contiguous(), student_masks.float().contiguous()
        B = pm.shape[0]
        self.flat_g.zero_()
        self.loss.zero_()
        prob_p, w, ws_p = self._forward(pm)           # the reference evaluates the pseudo labels first (:244-245)
        self._backward(pm, prob_p, w, ws_p, 1.0, 2 * B)
        prob_s, w, ws_s = self._forward(sm)
        self._backward(sm, prob_s, w, ws_s, 0.0, 2 * B)
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(self.pg)
            if world > 1:
                torch.distributed.all_reduce(self.flat_g, group=self.pg)
        self.opt_steps += 1
        lr = self.lr0 * self.gamma ** ((self.opt_steps - 1) // self.step_size)
        dev = self.flat_p.device
        with torch.cuda.device(dev):
            _lib.call("ucod_adamw_ema_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), ptr(self.flat_v),
                      ptr(None), _u64(self.n), c_float(lr), c_float(self.betas[0]), c_float(self.betas[1]),
                      c_float(self.eps), c_float(self.wd), int(self.opt_steps), c_float(1.0 / world), c_float(0.0),
                      stream_ptr(dev))
        self.last = {"probs_pseudo": prob_p, "probs_student": prob_s}
        return self.loss.clone()

    @torch.no_grad()
    def epoch_step(self, model, key_tokens_bf16: torch.Tensor, grid_in, pseudo_labels: torch.Tensor, feature_size=68):
        """One `Discriminator_epoch` iteration from cached keys: student forward (no grad) -> sigmoid > 0.5,
        pseudo labels bilinear 16 -> fs then > 0.5, then `step`."""
        fs = (feature_size, feature_size)
        fg, _, _ = model.decoder.forward_tokens(key_tokens_bf16, grid_in, fs, want_bg=False)
        pl = ops.upsample_bilinear(pseudo_labels.float(), fs)
        s_mask, _, p_mask = ops.apm_binarize(fg, fg, pl)
        return self.step(p_mask, s_mask)
