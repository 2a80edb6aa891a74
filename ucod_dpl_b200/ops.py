"""Thin Python wrappers over the C-ABI kernels (one function per `ucod_*` entry point).

All functions take CUDA tensors, allocate outputs with torch (device memory plumbing only) and launch on the
current CUDA stream.  No CPU fallback exists: non-CUDA inputs raise `UcodError`.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import UcodError, c_float, c_int, ptr, stream_ptr

_u64 = ctypes.c_uint64
_i64 = ctypes.c_int64


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=device)


def _aligned(ws: torch.Tensor):
    off = (-ws.data_ptr()) % 1024
    return ctypes.c_void_p(ws.data_ptr() + off), _u64(ws.numel() - off)


# ------------------------------------------------------------------------------------------------
def features_to_tokens_bf16(features: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 (any strides for C / HW as long as HW is jointly strided) -> [B,H*W,C] bf16."""
    _lib.require_cuda(features)
    if features.dtype != torch.float32:
        features = features.float()
    B, C, H, W = features.shape
    if features.stride(2) != W * features.stride(3):
        features = features.contiguous()
    out = torch.empty(B, H * W, C, device=features.device, dtype=torch.bfloat16)
    with torch.cuda.device(features.device):
        _lib.call("ucod_features_to_tokens_bf16", ptr(features), ptr(out), B, C, H * W, _i64(features.stride(0)),
                  _i64(features.stride(1)), _i64(features.stride(3)), stream_ptr(features.device))
    return out


def decoder_forward(keys_bf16: torch.Tensor, grid_in, grid_out, w_dec_bf16, b_dec, emb, w_fg, b_fg, w_bg, b_bg, *,
                    want_bg: bool = True, want_ortho: bool = False, count_dev: torch.Tensor | None = None):
    """keys_bf16 [B, gin_h*gin_w, dim] -> (fg [B,1,oh,ow], bg | None, ortho scalar tensor | None).
    count_dev: optional int32 device scalar, number of leading images to process (eval only, read on the device)."""
    _lib.require_cuda(keys_bf16)
    if keys_bf16.dtype != torch.bfloat16 or not keys_bf16.is_contiguous():
        raise UcodError("decoder_forward expects contiguous bf16 token-major keys")
    B, P, dim = keys_bf16.shape
    gh, gw = grid_in
    oh, ow = grid_out
    if gh * gw != P:
        raise UcodError(f"decoder_forward: {P} tokens do not form a {gh}x{gw} grid")
    dev = keys_bf16.device
    fg = torch.empty(B, 1, oh, ow, device=dev, dtype=torch.float32)
    bg = torch.empty(B, 1, oh, ow, device=dev, dtype=torch.float32) if want_bg else None
    ortho = torch.empty((), device=dev, dtype=torch.float32) if want_ortho else None
    lib = _lib.load()
    lib.ucod_decoder_workspace_bytes.restype = _u64
    need = lib.ucod_decoder_workspace_bytes(B, gh, gw, oh, ow, 1 if want_ortho else 0)
    ws = _ws(need, dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        if count_dev is None:
            _lib.call("ucod_decoder_fwd", ptr(keys_bf16), B, dim, gh, gw, oh, ow, ptr(w_dec_bf16), ptr(b_dec),
                      ptr(emb), ptr(w_fg), ptr(b_fg), ptr(w_bg), ptr(b_bg), ptr(fg), ptr(bg), ptr(ortho), wp, wn,
                      stream_ptr(dev))
        else:
            if want_ortho:
                raise UcodError("decoder_forward: count_dev is an eval-only option (no orthogonality loss)")
            _lib.call("ucod_decoder_fwd_dyn", ptr(keys_bf16), B, ptr(count_dev), dim, gh, gw, oh, ow, ptr(w_dec_bf16),
                      ptr(b_dec), ptr(emb), ptr(w_fg), ptr(b_fg), ptr(w_bg), ptr(b_bg), ptr(fg), ptr(bg), wp, wn,
                      stream_ptr(dev))
    return fg, bg, ortho


def upsample_bilinear(x: torch.Tensor, size, binarize=False) -> torch.Tensor:
    """F.interpolate(x, size, mode='bilinear', align_corners=False) for [B,1,h,w] / [B,h,w] fp32 maps;
    `binarize=True|1` returns the uint8 mask sigmoid(up(x)) > 0.5, 2: up(sigmoid(x)) > 0.5, 3: up(x) > 0.5."""
    _lib.require_cuda(x)
    shp = x.shape
    x3 = x.reshape(-1, shp[-2], shp[-1]).float().contiguous()
    oh, ow = size
    out = torch.empty(x3.shape[0], oh, ow, device=x.device, dtype=torch.uint8 if binarize else torch.float32)
    with torch.cuda.device(x.device):
        _lib.call("ucod_upsample_bilinear", ptr(x3), ptr(out), x3.shape[0], shp[-2], shp[-1], oh, ow,
                  int(binarize), stream_ptr(x.device))
    return out.reshape(*shp[:-2], oh, ow)


# ------------------------------------------------------------------------------------------------
def pseudo_label_score(attn_cls: torch.Tensor, keys: torch.Tensor, th_bkg: float, epsilon: float = 1e-10,
                       want_sim: bool = False, apply_weights: bool = True):
    """attn_cls [B,heads,P] fp32, keys [B,P,heads*64] fp32|bf16 -> (cos [B,P], bkg u8 [B,P], ref_idx [B], sim|None)."""
    _lib.require_cuda(attn_cls, keys)
    attn_cls = attn_cls.float().contiguous()
    keys = keys.contiguous()
    if keys.dtype not in (torch.float32, torch.bfloat16):
        keys = keys.float()
    B, nh, P = attn_cls.shape
    if keys.shape != (B, P, nh * 64):
        raise UcodError(f"pseudo_label_score: keys {tuple(keys.shape)} do not match attention {tuple(attn_cls.shape)}")
    dev = attn_cls.device
    cos = torch.empty(B, P, device=dev, dtype=torch.float32)
    bkg = torch.empty(B, P, device=dev, dtype=torch.uint8)
    ref = torch.empty(B, device=dev, dtype=torch.int32)
    sim = torch.empty(B, P, device=dev, dtype=torch.float32) if want_sim else None
    lib = _lib.load()
    lib.ucod_pseudo_label_scratch_bytes.restype = _u64
    ws = _ws(lib.ucod_pseudo_label_scratch_bytes(B, nh), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_pseudo_label_score_ex", ptr(attn_cls), ptr(keys), 1 if keys.dtype == torch.bfloat16 else 0, B,
                  nh, P, c_float(th_bkg), c_float(epsilon), 1 if apply_weights else 0, ptr(cos), ptr(bkg), ptr(ref),
                  ptr(sim), wp, wn, stream_ptr(dev))
    return cos, bkg, ref, sim


def refine_small_components(mask_u8: torch.Tensor, area_threshold: int = 4) -> torch.Tensor:
    """mask uint8 {0,1} [B,h,w] -> refined uint8 [B,h,w] (refine_post_process, batched)."""
    _lib.require_cuda(mask_u8)
    m = mask_u8.to(torch.uint8).contiguous()
    B, h, w = m.shape
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _lib.call("ucod_refine_small_components", ptr(m), ptr(out), B, h, w, int(area_threshold), stream_ptr(m.device))
    return out


# ------------------------------------------------------------------------------------------------
LT_MAX_BOXES = 128


LT_ALGORITHMS = {"auto": 0, "global": 1, "shared": 2}


def lt_boxes(mask_u8: torch.Tensor, look_twice_th: float, expand_type: str = "dynamic", const_scale: float = 1.3,
             want_labels: bool = False, algorithm: str = "auto"):
    """mask uint8 [B,H,W] -> (boxes int32 [B,128,4], nbox int32 [B], status int32 [B], labels int32 [B,H,W] | None).
    algorithm: "auto" / "shared" label each mask in one CTA's shared memory (run-based; nbox = -3 for a mask that
    exceeds its capacity — callers re-run with "global"), "global" is the union-find over pixels in HBM."""
    _lib.require_cuda(mask_u8)
    m = mask_u8.to(torch.uint8).contiguous()
    B, H, W = m.shape
    dev = m.device
    boxes = torch.zeros(B, LT_MAX_BOXES, 4, device=dev, dtype=torch.int32)
    nbox = torch.empty(B, device=dev, dtype=torch.int32)
    status = torch.empty(B, device=dev, dtype=torch.int32)
    labels = torch.empty(B, H, W, device=dev, dtype=torch.int32) if want_labels else None
    lib = _lib.load()
    lib.ucod_lt_boxes_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_lt_boxes_workspace_bytes(B, H, W), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_lt_boxes_ex", ptr(m), B, H, W, ctypes.c_double(look_twice_th),
                  1 if expand_type == "dynamic" else 0, ctypes.c_double(const_scale), ptr(boxes), ptr(nbox),
                  ptr(status), ptr(labels), wp, wn, LT_ALGORITHMS[algorithm], stream_ptr(dev))
    return boxes, nbox, status, labels


def roi_crop_resize(images_u8: torch.Tensor, jobs: torch.Tensor, out_size, layout: str = "CHW") -> torch.Tensor:
    """images uint8 [N,3,H0,W0] (layout 'CHW') or [N,H0,W0,3] ('HWC'); jobs int32 [n,5] (img,x,y,w,h) on the same
    device -> uint8 [n,3,out_h,out_w] = PIL crop + antialiased bilinear Resize."""
    _lib.require_cuda(images_u8, jobs)
    if images_u8.dtype != torch.uint8:
        raise UcodError("roi_crop_resize expects uint8 images")
    if layout == "CHW":
        N, _, H0, W0 = images_u8.shape
        s_img, s_ch, s_row, s_px = images_u8.stride()
    else:
        N, H0, W0, _ = images_u8.shape
        s_img, s_row, s_px, s_ch = images_u8.stride()
    jobs = jobs.to(torch.int32).contiguous()
    n = jobs.shape[0]
    oh, ow = out_size
    dev = images_u8.device
    out = torch.empty(n, 3, oh, ow, device=dev, dtype=torch.uint8)
    if n == 0:
        return out
    max_h = max(int(jobs[:, 4].max().item()), 1)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    lib = _lib.load()
    lib.ucod_roi_crop_resize_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_roi_crop_resize_workspace_bytes(n, max_h, oh, ow), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_roi_crop_resize", ptr(images_u8), N, H0, W0, _i64(s_img), _i64(s_ch), _i64(s_row), _i64(s_px),
                  ptr(jobs), n, max_h, ptr(out), oh, ow, wp, wn, ptr(err), stream_ptr(dev))
    if int(err.item()) & 1:  # (this wrapper already synchronised for max_h; the *_dyn path defers the check)
        raise UcodError("roi_crop_resize: a crop is down-scaled by more than ~19x, beyond this call's tap table; "
                        "use roi_crop_resize_dyn (no limit)")
    return out


def lt_build_jobs(boxes: torch.Tensor, nbox: torch.Tensor, mask_size, src_size=None, orig_sizes=None, *, capacity: int,
                  chunk: int):
    """Device-side Look-Twice job tables (no host synchronisation): boxes int32 [B,128,4], nbox int32 [B] from
    `lt_boxes` -> (crop_jobs int32 [capacity,5], paste_jobs int32 [capacity,6], counts int32 [4] =
    (jobs kept, status bits, jobs requested, 0), chunk_counts int32 [ceil(capacity / chunk)])."""
    _lib.require_cuda(boxes, nbox)
    B = nbox.shape[0]
    dev = boxes.device
    sh, sw = mask_size
    if orig_sizes is not None:
        orig_sizes = orig_sizes.to(device=dev, dtype=torch.int32).contiguous()
        h0 = w0 = 0
    else:
        h0, w0 = src_size
    crop = torch.zeros(capacity, 5, device=dev, dtype=torch.int32)
    paste = torch.zeros(capacity, 6, device=dev, dtype=torch.int32)
    counts = torch.zeros(4, device=dev, dtype=torch.int32)
    chunks = torch.zeros((capacity + chunk - 1) // chunk, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.call("ucod_lt_build_jobs", ptr(boxes), ptr(nbox), B, sh, sw, int(h0), int(w0), ptr(orig_sizes), ptr(crop),
                  ptr(paste), capacity, ptr(counts), chunk, ptr(chunks), stream_ptr(dev))
    return crop, paste, counts, chunks


def roi_crop_resize_dyn(images_u8: torch.Tensor, jobs: torch.Tensor, njobs_dev: torch.Tensor, out_size,
                        layout: str = "CHW", err: torch.Tensor | None = None) -> torch.Tensor:
    """`roi_crop_resize` for a job slice whose length is a device scalar (`njobs_dev`, int32, <= jobs.shape[0]).
    No host synchronisation; `err` (int32 [1], accumulated, bit 0 = tap table overflow) is the caller's to check."""
    _lib.require_cuda(images_u8, jobs, njobs_dev)
    if images_u8.dtype != torch.uint8 or jobs.dtype != torch.int32 or not jobs.is_contiguous():
        raise UcodError("roi_crop_resize_dyn expects uint8 images and a contiguous int32 job table")
    if layout == "CHW":
        N, _, H0, W0 = images_u8.shape
        s_img, s_ch, s_row, s_px = images_u8.stride()
    else:
        N, H0, W0, _ = images_u8.shape
        s_img, s_row, s_px, s_ch = images_u8.stride()
    cap = jobs.shape[0]
    oh, ow = out_size
    dev = images_u8.device
    out = torch.empty(cap, 3, oh, ow, device=dev, dtype=torch.uint8)
    if err is None:
        err = torch.zeros(1, device=dev, dtype=torch.int32)
    lib = _lib.load()
    lib.ucod_roi_crop_resize_dyn_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_roi_crop_resize_dyn_workspace_bytes(cap, H0, W0, oh, ow), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_roi_crop_resize_dyn", ptr(images_u8), N, H0, W0, _i64(s_img), _i64(s_ch), _i64(s_row),
                  _i64(s_px), ptr(jobs), cap, ptr(njobs_dev), ptr(out), oh, ow, wp, wn, ptr(err), stream_ptr(dev))
    return out


def paste_bicubic_dyn(logits: torch.Tensor, all_jobs: torch.Tensor, first_index: int, njobs_dev: torch.Tensor,
                      n_all_dev: torch.Tensor, mask_u8: torch.Tensor, out_cap: int,
                      err: torch.Tensor | None = None) -> None:
    """Paste the slice [first_index, first_index + njobs_dev) of an image-major, rank-ascending paste table
    (`lt_build_jobs`) into mask uint8 [N,S,S]; logits fp32 [capacity,g,g] are the slice's second-look predictions.
    A pixel takes the value of the last job of its image that covers it, as the reference's sequential pastes do."""
    _lib.require_cuda(logits, all_jobs, njobs_dev, n_all_dev, mask_u8)
    logits = logits.float().contiguous()
    cap, gh, gw = logits.shape
    N, Sh, Sw = mask_u8.shape
    dev = logits.device
    if err is None:
        err = torch.zeros(1, device=dev, dtype=torch.int32)
    lib = _lib.load()
    lib.ucod_paste_bicubic_dyn_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_paste_bicubic_dyn_workspace_bytes(cap, gh, gw, out_cap), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_paste_bicubic_dyn", ptr(logits), cap, ptr(njobs_dev), gh, gw, ptr(all_jobs), int(first_index),
                  all_jobs.shape[0], ptr(n_all_dev), ptr(mask_u8), N, Sh, Sw, int(out_cap), wp, wn, ptr(err),
                  stream_ptr(dev))


def paste_bicubic(logits: torch.Tensor, jobs: torch.Tensor, mask_u8: torch.Tensor) -> None:
    """logits fp32 [n,g,g]; jobs int32 [n,6] (img,x,y,w,h,rank); mask uint8 [N,S,S] with values 0..255, updated
    in place (binarise -> PIL bicubic resize -> paste, rank order per image)."""
    _lib.require_cuda(logits, jobs, mask_u8)
    logits = logits.float().contiguous()
    jobs = jobs.to(torch.int32).contiguous()
    n, gh, gw = logits.shape
    if n == 0:
        return
    N, Sh, Sw = mask_u8.shape
    cap = max(int(jobs[:, 3:5].max().item()), 1)
    max_rank = int(jobs[:, 5].max().item())
    dev = logits.device
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    lib = _lib.load()
    lib.ucod_paste_bicubic_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_paste_bicubic_workspace_bytes(n, gh, cap), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_paste_bicubic", ptr(logits), n, gh, gw, ptr(jobs), max_rank, ptr(mask_u8), N, Sh, Sw, cap, wp,
                  wn, ptr(err), stream_ptr(dev))
    if int(err.item()) & 1:
        raise UcodError("paste_bicubic: a box is smaller than ~4 px, beyond this call's tap table; use paste_bicubic_dyn")


def mask_scale_u8(mask_u8: torch.Tensor, mul: int = 255) -> torch.Tensor:
    _lib.require_cuda(mask_u8)
    m = mask_u8.contiguous()
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _lib.call("ucod_mask_scale_u8", ptr(m), ptr(out), _u64(m.numel()), int(mul), stream_ptr(m.device))
    return out


def to_tensor_normalize(images_u8: torch.Tensor, mean=None, std=None) -> torch.Tensor:
    """planar uint8 [..., C, H, W] -> fp32 of the same shape: torchvision ToTensor (+ Normalize when `mean`/`std`
    are given), bit-exact (data/datasets/transforms.py:14-18)."""
    _lib.require_cuda(images_u8)
    if images_u8.dtype != torch.uint8 or images_u8.dim() < 3:
        raise UcodError("to_tensor_normalize expects a planar uint8 tensor [..., C, H, W]")
    x = images_u8.contiguous()
    C, H, W = x.shape[-3:]
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    if mean is None:
        ch, m, sd = 0, None, None
    else:
        if len(mean) != C or len(std) != C or C > 3:
            raise UcodError("to_tensor_normalize: mean/std must have one entry per channel (<= 3 channels)")
        ch = C
        m = (ctypes.c_float * C)(*[float(v) for v in mean])
        sd = (ctypes.c_float * C)(*[float(v) for v in std])
    with torch.cuda.device(x.device):
        _lib.call("ucod_to_tensor_normalize", ptr(x), ptr(out), _u64(x.numel() // (H * W)), H * W, ch, m, sd,
                  stream_ptr(x.device))
    return out


# ------------------------------------------------------------------------------------------------
class _DiscW(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "conv1", "bn1_w", "bn1_b", "bn1_mean", "bn1_var", "conv2", "bn2_w", "bn2_b", "bn2_mean", "bn2_var",
        "conv3", "bn3_w", "bn3_b", "bn3_mean", "bn3_var", "lin_w", "lin_b")]


def discriminator_forward(mask: torch.Tensor, tensors: dict, bn_train: bool = True,
                          update_running: bool = True) -> torch.Tensor:
    """mask fp32 [B,1,fs,fs]; `tensors`: fp32 CUDA tensors keyed by the _DiscW field names -> prob [B,1]."""
    _lib.require_cuda(mask)
    m = mask.float().contiguous()
    B, _, fs, fs2 = m.shape
    if fs != fs2:
        raise UcodError("discriminator_forward expects square masks")
    dev = m.device
    w = _DiscW(*[tensors[n].data_ptr() for n, _ in _DiscW._fields_])
    prob = torch.empty(B, 1, device=dev, dtype=torch.float32)
    lib = _lib.load()
    lib.ucod_discriminator_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_discriminator_workspace_bytes(B, fs), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_discriminator_fwd", ptr(m), B, fs, ctypes.byref(w), 1 if bn_train else 0,
                  1 if update_running else 0, ptr(prob), wp, wn, stream_ptr(dev))
    return prob


def discriminator_forward_calls(masks: torch.Tensor, calls: int, tensors: dict, bn_train: bool = True,
                                update_running: bool = True) -> torch.Tensor:
    """`calls` consecutive discriminator forwards in one set of launches: masks fp32 [calls*B,1,fs,fs] (call-major)
    -> prob [calls*B,1]; batch statistics per call, running buffers updated in call order."""
    _lib.require_cuda(masks)
    m = masks.float().contiguous()
    n, _, fs, fs2 = m.shape
    if fs != fs2 or n % calls:
        raise UcodError("discriminator_forward_calls expects square masks, calls * B of them")
    B, dev = n // calls, m.device
    w = _DiscW(*[tensors[k].data_ptr() for k, _ in _DiscW._fields_])
    prob = torch.empty(n, 1, device=dev, dtype=torch.float32)
    lib = _lib.load()
    lib.ucod_discriminator_workspace_bytes_calls.restype = _u64
    ws = _ws(lib.ucod_discriminator_workspace_bytes_calls(B, fs, calls), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_discriminator_fwd_calls", ptr(m), B, calls, fs, ctypes.byref(w), 1 if bn_train else 0,
                  1 if update_running else 0, ptr(prob), wp, wn, stream_ptr(dev))
    return prob


def apm_binarize(student: torch.Tensor, teacher: torch.Tensor, pl: torch.Tensor):
    """-> (s_mask, t_mask, p_mask); s_mask and p_mask are the two halves of one [2B,...] buffer (`.pair`) so that both
    discriminator calls of the APM can share a launch."""
    _lib.require_cuda(student, teacher, pl)
    s, t, p = student.float().contiguous(), teacher.float().contiguous(), pl.float().contiguous()
    pair = torch.empty((2 * s.shape[0],) + tuple(s.shape[1:]), device=s.device, dtype=torch.float32)
    sm, pm = pair[:s.shape[0]], pair[s.shape[0]:]
    tm = torch.empty_like(t)
    sm.pair = pair
    with torch.cuda.device(s.device):
        _lib.call("ucod_apm_binarize", ptr(s), ptr(t), ptr(p), ptr(sm), ptr(tm), ptr(pm), _u64(s.numel()),
                  stream_ptr(s.device))
    return sm, tm, pm


def apm_merge(pl: torch.Tensor, t_mask: torch.Tensor, p_s: torch.Tensor, p_p: torch.Tensor, epoch_term: float):
    """-> (merged like pl, weight [B,1], dis_loss scalar tensor)."""
    _lib.require_cuda(pl, t_mask, p_s, p_p)
    pl = pl.float().contiguous()
    B = pl.shape[0]
    npix = pl.numel() // B
    merged = torch.empty_like(pl)
    weight = torch.empty(B, 1, device=pl.device, dtype=torch.float32)
    loss = torch.empty((), device=pl.device, dtype=torch.float32)
    with torch.cuda.device(pl.device):
        _lib.call("ucod_apm_merge", ptr(pl), ptr(t_mask.contiguous()), ptr(p_s.float().contiguous()),
                  ptr(p_p.float().contiguous()), c_float(epoch_term), ptr(merged), ptr(weight), ptr(loss), B, npix,
                  stream_ptr(pl.device))
    return merged, weight, loss


# ------------------------------------------------------------------------------------------------
# dense building blocks (host-assembled blocks: CORAL CSF)
def gemm_bf16(a: torch.Tensor, w: torch.Tensor, mode: int, bias: torch.Tensor | None = None,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """out = epilogue(a @ w.T): a [M,K] bf16, w [N,K] bf16 (row pitches may exceed K).  mode 0: bf16 (+bias),
    1: bf16 gelu(+bias), 2: fp32 `out += acc + bias` (out required), 5: fp32 (+bias)."""
    _lib.require_cuda(a, w)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        if mode == 2:
            raise UcodError("gemm_bf16 mode 2 accumulates into `out`")
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if mode == 5 else torch.bfloat16)
    with torch.cuda.device(a.device):
        _lib.call("ucod_gemm_bf16", ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, int(mode), ptr(bias), ptr(out),
                  out.stride(0), stream_ptr(a.device))
    return out


def layernorm_bf16(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    """x fp32 [rows, dim] contiguous -> bf16 [rows, dim]."""
    _lib.require_cuda(x)
    rows, dim = x.shape
    y = torch.empty(rows, dim, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.call("ucod_layernorm_bf16", ptr(x), ptr(weight), ptr(bias), ptr(y), rows, dim, c_float(eps),
                  stream_ptr(x.device))
    return y


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(x)
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.call("ucod_cast_f32_bf16", ptr(x), ptr(y), _u64(x.numel()), stream_ptr(x.device))
    return y


def features_to_tokens_f32(features: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 -> token-major fp32 [B,H*W,C]."""
    _lib.require_cuda(features)
    features = features.float()
    B, C, H, W = features.shape
    if features.stride(2) != W * features.stride(3):
        features = features.contiguous()
    out = torch.empty(B, H * W, C, device=features.device, dtype=torch.float32)
    with torch.cuda.device(features.device):
        _lib.call("ucod_features_to_tokens_f32", ptr(features), ptr(out), B, C, H * W, _i64(features.stride(0)),
                  _i64(features.stride(1)), _i64(features.stride(3)), stream_ptr(features.device))
    return out


def resize_tokens_bilinear(tokens: torch.Tensor, grid_in, grid_out, want_f32: bool = True, want_bf16: bool = False):
    """tokens fp32 [n, gh*gw, C] -> (fp32 | None, bf16 | None) on the `grid_out` grid (bilinear, align_corners=False)."""
    _lib.require_cuda(tokens)
    tokens = tokens.float().contiguous()
    n, P, C = tokens.shape
    (gh, gw), (oh, ow) = grid_in, grid_out
    if gh * gw != P:
        raise UcodError("resize_tokens_bilinear: token count does not match the input grid")
    o32 = torch.empty(n, oh * ow, C, device=tokens.device, dtype=torch.float32) if want_f32 else None
    o16 = torch.empty(n, oh * ow, C, device=tokens.device, dtype=torch.bfloat16) if want_bf16 else None
    with torch.cuda.device(tokens.device):
        _lib.call("ucod_resize_tokens_bilinear", ptr(tokens), ptr(o32), ptr(o16), n, gh, gw, oh, ow, C,
                  stream_ptr(tokens.device))
    return o32, o16


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, head_dim: int, scale: float,
              kv_batch_map: torch.Tensor | None = None, head_dim_real: int | None = None) -> torch.Tensor:
    """q [B,Tq,>=heads*head_dim] bf16, k / v [Bkv,Tk,...] bf16 views with the same row pitch -> ctx bf16
    [B,Tq,heads*head_dim].  `kv_batch_map` int32 [B]: K/V batch of each q batch."""
    _lib.require_cuda(q, k, v)
    B, Tq = q.shape[0], q.shape[1]
    Bkv, Tk = k.shape[0], k.shape[1]
    if k.stride(1) != v.stride(1) or q.stride(2) != 1 or k.stride(2) != 1 or v.stride(2) != 1:
        raise UcodError("attention: k and v must share a row pitch and be unit-stride in the last dim")
    if q.stride(0) != Tq * q.stride(1) or k.stride(0) != Tk * k.stride(1):
        raise UcodError("attention: batch stride must equal tokens * row pitch")
    ctx = torch.empty(B, Tq, heads * head_dim, device=q.device, dtype=torch.bfloat16)
    with torch.cuda.device(q.device):
        _lib.call("ucod_attention_shared_kv", ptr(q), q.stride(1), ptr(k), ptr(v), k.stride(1), ptr(ctx),
                  heads * head_dim, B, heads, head_dim, int(head_dim_real or head_dim), Tq, Tk, c_float(scale),
                  ptr(kv_batch_map), Bkv if kv_batch_map is not None else 0, stream_ptr(q.device))
    return ctx


# ------------------------------------------------------------------------------------------------
# CORAL second stage
def coral_entropy_select(preds: torch.Tensor, threshold: float, window_size: int, per_image: bool = False):
    """preds [B,1,P,P] -> (entropy [B,1,P,P], scores [B,1,w,w], mask bool [B,1,w,w]).
    per_image: decide "probabilities or logits" (ASR.py:42) for every image on its own, as the reference's batch-1 eval
    loop does, instead of once for the whole call."""
    _lib.require_cuda(preds)
    p = preds.float().contiguous()
    B, _, P, _ = p.shape
    dev = p.device
    entropy = torch.empty_like(p)
    scores = torch.empty(B, 1, window_size, window_size, device=dev, dtype=torch.float32)
    mask = torch.empty(B, 1, window_size, window_size, device=dev, dtype=torch.uint8)
    scratch = torch.empty(B, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.call("ucod_coral_entropy_select_ex", ptr(p), B, P, window_size, c_float(threshold), ptr(entropy),
                  ptr(scores), ptr(mask), ptr(scratch), 1 if per_image else 0, stream_ptr(dev))
    return entropy, scores, mask.bool()


def coral_window_head(taps: torch.Tensor, n_windows: int, grid: int, bias_const: float) -> torch.Tensor:
    _lib.require_cuda(taps)
    out = torch.empty(n_windows, 1, grid, grid, device=taps.device, dtype=torch.float32)
    with torch.cuda.device(taps.device):
        _lib.call("ucod_coral_window_head", ptr(taps), taps.stride(0), n_windows, grid, c_float(bias_const), ptr(out),
                  stream_ptr(taps.device))
    return out


def coral_scatter_windows(window_preds: torch.Tensor | None, slot_of_cell: torch.Tensor, batch: int, window_size: int,
                          grid: int) -> torch.Tensor:
    _lib.require_cuda(slot_of_cell)
    dev = slot_of_cell.device
    S = window_size * grid
    out = torch.empty(batch, 1, S, S, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.call("ucod_coral_scatter_windows", ptr(window_preds), ptr(slot_of_cell), batch, window_size, grid,
                  ptr(out), stream_ptr(dev))
    return out


def coral_gated_ensemble(preds: torch.Tensor, h_preds: torch.Tensor, w0, b0, w2, b2, max_per_image: bool = False):
    """preds [B,1,P,P], h_preds [B,1,S,S] -> (out [B,1,S,S], weight [B,1,S,S]).  max_per_image: normalise the local
    entropy by each image's own maximum (= the reference run at batch 1) instead of the batch maximum."""
    _lib.require_cuda(preds, h_preds)
    p, h = preds.float().contiguous(), h_preds.float().contiguous()
    B, _, P, _ = p.shape
    S = h.shape[-1]
    dev = p.device
    out, weight = torch.empty_like(h), torch.empty_like(h)
    lib = _lib.load()
    lib.ucod_coral_gated_ensemble_workspace_bytes.restype = _u64
    ws = _ws(lib.ucod_coral_gated_ensemble_workspace_bytes(B, S), dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_coral_gated_ensemble", ptr(p), P, ptr(h), B, S, int(bool(max_per_image)), ptr(w0), ptr(b0),
                  ptr(w2), ptr(b2), ptr(out), ptr(weight), wp, wn, stream_ptr(dev))
    return out, weight
