"""CORAL refiner modules with the reference's class names, constructors and parameter names
(models/modules/{ASR,HRE,CSF,GE_pix_level,mlp}.py): `EntropySelector`, `CrossAttentionBlock`, `CSF`, `HRE`,
`GatedEnsembler`.  The torch modules are parameter containers; every forward runs in the CUDA library:

  CSF.forward  = LayerNorm(q), LayerNorm(kv) -> in_proj GEMMs -> tcgen05 attention (8 heads x 96, zero-padded to
                 128; K/V computed once per image and shared by its selected windows) -> out_proj GEMM with the
                 residual reduce-add -> LayerNorm -> fc1+GELU GEMM -> fc2 GEMM (+residual) -> [dw-conv 7x7 o 1x1
                 mask_dec] folded into a 768->49 tap GEMM + 49-tap gather-sum.
"""
from __future__ import annotations

import torch
from torch import nn

from ... import ops

HEADS, HEAD_DIM, HEAD_PAD = 8, 96, 128


class EntropySelector(nn.Module):
    """models/modules/ASR.py:7-51.  Returns the same tuple as the reference:
    (l_input_features [N,C,g,g], h_inputs_features [N,C,g,g], mask bool [B,1,w,w], coords_list [N,2], entropy)."""

    def __init__(self, threshold: float, window_size: int) -> None:
        super().__init__()
        self.threshold = threshold
        self.window_size = window_size

    @torch.no_grad()
    def select(self, preds, per_image: bool = False):
        """-> (mask bool [B,1,w,w], entropy [B,1,P,P], window->image index list, coords [N,2] CPU)."""
        entropy, _, mask = ops.coral_entropy_select(preds, self.threshold, self.window_size, per_image=per_image)
        flat = mask.flatten(1).cpu()  # the one host round trip of the refiner: the number of selected windows
        win_img, coords = [], []
        for b in range(flat.shape[0]):
            for cell in torch.nonzero(flat[b]).flatten().tolist():
                win_img.append(b)
                coords.append([cell // self.window_size, cell % self.window_size])
        return mask, entropy, win_img, torch.tensor(coords, dtype=torch.long).reshape(-1, 2), flat

    def forward(self, input_features, h_inputs, preds):
        mask, entropy, win_img, coords, flat = self.select(preds)
        idx = torch.nonzero(flat.flatten()).flatten().to(h_inputs.device)
        h_sel = h_inputs.flatten(0, 1).index_select(0, idx)
        l_rep = input_features.index_select(0, torch.tensor(win_img, dtype=torch.long, device=input_features.device))
        return l_rep, h_sel, mask, coords.to(preds.device), entropy


class CrossAttentionBlock(nn.Module):
    """models/modules/mlp.py:116-148 (parameter container + packed-weight cache)."""

    def __init__(self, dim, num_heads=8, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim != HEADS * HEAD_DIM or num_heads != HEADS:
            raise NotImplementedError("the CUDA CSF block is built for dim 768 = 8 heads x 96")
        self.norm_q = norm_layer(dim)
        self.norm_kv = norm_layer(dim)
        self.attn = nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, dropout=attn_drop, batch_first=True)
        self.drop_path = nn.Identity()
        hidden = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), act_layer(), nn.Linear(hidden, dim), nn.Dropout(drop))
        self.norm_mlp = norm_layer(dim)


def _pad_heads_rows(w: torch.Tensor) -> torch.Tensor:
    """[8*96, ...] -> [8*128, ...] with zero rows after each head's 96."""
    out = w.new_zeros((HEADS * HEAD_PAD,) + tuple(w.shape[1:]))
    for h in range(HEADS):
        out[h * HEAD_PAD:h * HEAD_PAD + HEAD_DIM] = w[h * HEAD_DIM:(h + 1) * HEAD_DIM]
    return out


class CSF(nn.Module):
    """models/modules/CSF.py:7-43."""

    def __init__(self, dim=768) -> None:
        super().__init__()
        self.dim = dim
        self.attn = CrossAttentionBlock(dim=dim)
        self.depthwise_conv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.mask_dec = nn.Conv2d(dim, 1, kernel_size=1, padding=0)
        self._packed = None

    def _pack(self):
        params = list(self.parameters())
        tag = tuple((p._version, p.data_ptr()) for p in params)
        if self._packed is not None and self._packed[0] == tag:
            return self._packed[1]
        C = self.dim
        a = self.attn
        w_in, b_in = a.attn.in_proj_weight.detach().float(), a.attn.in_proj_bias.detach().float()
        bf = lambda t: t.to(torch.bfloat16).contiguous()  # noqa: E731
        f32 = lambda t: t.detach().float().contiguous()   # noqa: E731
        w_o = a.attn.out_proj.weight.detach().float()
        w_o_pad = w_o.new_zeros(C, HEADS * HEAD_PAD)
        for h in range(HEADS):
            w_o_pad[:, h * HEAD_PAD:h * HEAD_PAD + HEAD_DIM] = w_o[:, h * HEAD_DIM:(h + 1) * HEAD_DIM]
        dw, md = self.depthwise_conv.weight.detach().float(), self.mask_dec.weight.detach().float().reshape(C)
        taps = (dw.reshape(C, 49) * md[:, None]).t()                      # [49, C]
        w_taps = taps.new_zeros(128, C)
        w_taps[:49] = taps
        const = float((md * self.depthwise_conv.bias.detach().float()).sum() + self.mask_dec.bias.detach().float()[0])
        pk = dict(
            w_q=bf(_pad_heads_rows(w_in[:C])), b_q=f32(_pad_heads_rows(b_in[:C])),
            w_kv=bf(torch.cat([_pad_heads_rows(w_in[C:2 * C]), _pad_heads_rows(w_in[2 * C:])], 0)),
            b_kv=f32(torch.cat([_pad_heads_rows(b_in[C:2 * C]), _pad_heads_rows(b_in[2 * C:])], 0)),
            w_o=bf(w_o_pad), b_o=f32(a.attn.out_proj.bias),
            w_1=bf(a.mlp[0].weight.detach().float()), b_1=f32(a.mlp[0].bias),
            w_2=bf(a.mlp[2].weight.detach().float()), b_2=f32(a.mlp[2].bias),
            w_taps=bf(w_taps), const=const,
            ln=[(f32(m.weight), f32(m.bias), m.eps) for m in (a.norm_q, a.norm_kv, a.norm_mlp)])
        self._packed = (tag, pk)
        return pk

    @torch.no_grad()
    def forward_tokens(self, l_tokens: torch.Tensor, h_tokens: torch.Tensor, win_img: torch.Tensor, grid: int):
        """l_tokens fp32 [B, g*g, C] (one per image), h_tokens fp32 [N, g*g, C] (selected windows),
        win_img int32 [N] (image of each window) -> window logits [N,1,g,g]."""
        pk = self._pack()
        N, T, C = h_tokens.shape
        B = l_tokens.shape[0]
        x = h_tokens.reshape(N * T, C).clone()                                       # fp32 residual stream
        (wq, bq, eq), (wkv, bkv, ekv), (wm, bm, em) = pk["ln"]
        qn = ops.layernorm_bf16(h_tokens.reshape(N * T, C), wq, bq, eq)
        kvn = ops.layernorm_bf16(l_tokens.reshape(B * T, C), wkv, bkv, ekv)
        Q = ops.gemm_bf16(qn, pk["w_q"], 0, pk["b_q"]).reshape(N, T, HEADS * HEAD_PAD)
        KV = ops.gemm_bf16(kvn, pk["w_kv"], 0, pk["b_kv"]).reshape(B, T, 2 * HEADS * HEAD_PAD)
        ctx = ops.attention(Q, KV[..., :HEADS * HEAD_PAD], KV[..., HEADS * HEAD_PAD:], HEADS, HEAD_PAD,
                            HEAD_DIM ** -0.5, kv_batch_map=win_img, head_dim_real=HEAD_DIM)
        ops.gemm_bf16(ctx.reshape(N * T, HEADS * HEAD_PAD), pk["w_o"], 2, pk["b_o"], out=x)   # x = query + attn
        hn = ops.layernorm_bf16(x, wm, bm, em)
        h1 = ops.gemm_bf16(hn, pk["w_1"], 1, pk["b_1"])
        ops.gemm_bf16(h1, pk["w_2"], 2, pk["b_2"], out=x)                                     # x += mlp
        taps = ops.gemm_bf16(ops.cast_bf16(x), pk["w_taps"], 5)
        return ops.coral_window_head(taps, N, grid, pk["const"])

    def forward(self, l_inputs, h_inputs):
        """Reference signature: l_inputs, h_inputs [N,C,g,g] (l already repeated per window) -> [N,1,g,g]."""
        N, C, g, _ = h_inputs.shape
        if N == 0:
            return torch.zeros(0, 1, g, g, device=h_inputs.device)
        win_img = torch.arange(N, dtype=torch.int32, device=h_inputs.device)
        return self.forward_tokens(ops.features_to_tokens_f32(l_inputs), ops.features_to_tokens_f32(h_inputs),
                                   win_img, g)


class HRE(nn.Module):
    """models/modules/HRE.py:7-44."""

    def __init__(self, window_size, dim=768) -> None:
        super().__init__()
        self.window_size = window_size
        self.CSF = CSF(dim)

    def concate_windows(self, windows, positions, candidate_windows_mask):
        B = candidate_windows_mask.shape[0]
        ws = self.window_size
        g = windows.shape[-1]
        flat = candidate_windows_mask.reshape(B, ws * ws).to(torch.int32)
        slot = (torch.cumsum(flat.flatten(), 0) - 1).to(torch.int32).reshape(B, ws * ws)
        slot = torch.where(flat > 0, slot, torch.full_like(slot, -1)).contiguous()
        return ops.coral_scatter_windows(windows.float().contiguous() if windows.numel() else None, slot, B, ws, g)

    def forward(self, l_input_features, h_inputs_features, candidate_windows_mask, coords_list):
        preds = self.CSF(l_input_features, h_inputs_features)
        return self.concate_windows(preds, coords_list, candidate_windows_mask), preds


class GatedEnsembler(nn.Module):
    """models/modules/GE_pix_level.py:6-26."""

    def __init__(self, num_classes: int) -> None:
        super().__init__()
        self.alpha = nn.Parameter(torch.tensor(0.5))
        self.fuser = nn.Sequential(nn.Conv2d(num_classes, 64, kernel_size=1), nn.ReLU(),
                                   nn.Conv2d(64, num_classes, kernel_size=1))

    def forward(self, l1: torch.Tensor, l2: torch.Tensor, max_per_image: bool = False):
        f0, f2 = self.fuser[0], self.fuser[2]
        return ops.coral_gated_ensemble(l1, l2, f0.weight.detach().reshape(64).float().contiguous(),
                                        f0.bias.detach().float().contiguous(),
                                        f2.weight.detach().reshape(64).float().contiguous(),
                                        f2.bias.detach().float().contiguous(), max_per_image=max_per_image)
