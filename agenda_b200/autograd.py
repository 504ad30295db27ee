"""Training-mode attention (SURVEY.md §8 f N3): the reference's only in-tree caller of hook.py runs the processor with
`is_train=True` under autograd (finetune_sd_token.py:754-757, 1043-1069): the fg/bg heat-map loss sends gradients
through the selected-token probabilities into Q and K of every cross-attention layer, and from there through every
self-attention layer of the UNet into the learned token embeddings.

The forward passes are the sm_100a kernels behind the C ABI (same entry points as inference).  Cross-attention backward
(where the heat-map gradient enters) is `agenda_attn_cross_bwd`, an exact fp32 CUDA-core kernel
(csrc/attn_cross_bwd.cu).  Self-attention backward is `agenda_attn_self_bwd` (csrc/attn_self_bwd_sm100.cu): tcgen05
kernels that recompute P from a log-sum-exp pass and form dQ, dK and dV with their accumulators in TMEM.

Gradient formulas (per batch element b and head h; P = softmax(scale * Q K^T), O = P V):
    dV = P^T dO
    dP = dO V^T  (+ dMaps[b', t, n] / H  on column token_idx[t], for b >= b_first: maps = mean over heads, hook.py:55)
    dS = P * (dP - rowsum(P * dP))
    dQ = scale * dS K,   dK = scale * dS^T Q
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

from . import ops


def _split_heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    B, S, C = x.shape
    return x.view(B, S, heads, C // heads).transpose(1, 2)  # [B, H, S, d]


def _merge_heads(x: torch.Tensor) -> torch.Tensor:
    B, H, S, d = x.shape
    return x.transpose(1, 2).reshape(B, S, H * d)


class SelfAttentionFn(torch.autograd.Function):
    """out = softmax(scale q k^T) v per head; q/k/v [B,N,H*d] (hook.py:104-115 with encoder_hidden_states None).

    Backward: precision="bf16" at the SD head dims -> the tcgen05 backward kernels (agenda_attn_self_bwd: log-sum-exp
    pass, then dQ / dK / dV with P recomputed on chip).  precision="fp32" (the exact-parity path) and other head dims ->
    the same formulas with fp32 GEMMs, one head at a time so that the [N, N] probability block stays bounded."""

    @staticmethod
    def forward(ctx, q, k, v, heads: int, scale: float, precision: str):
        lse = None
        if precision == "bf16" and (q.shape[-1] // heads) in ops.SELF_BWD_HEAD_DIMS:
            # the forward kernel's softmax epilogue also leaves every row's log-sum-exp: the backward then has no LSE pass
            out, lse = ops.attn_self_with_lse(q, k, v, heads, scale=scale)
        else:
            out = ops.attn_self(q, k, v, heads, scale=scale, precision=precision)
        ctx.save_for_backward(q, k, v, out, *(() if lse is None else (lse,)))
        ctx.heads, ctx.scale, ctx.precision = heads, scale, precision
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, out = ctx.saved_tensors[:4]
        lse = ctx.saved_tensors[4] if len(ctx.saved_tensors) > 4 else None
        d = q.shape[-1] // ctx.heads
        if ctx.precision == "bf16" and d in ops.SELF_BWD_HEAD_DIMS:
            dq, dk, dv = ops.attn_self_bwd(q, k, v, out, d_out, ctx.heads, scale=ctx.scale, lse=lse)
            return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None
        return SelfAttentionFn._backward_fp32(q, k, v, d_out, ctx.heads, ctx.scale) + (None, None, None)

    @staticmethod
    def _backward_fp32(q, k, v, d_out, heads, scale):
        qf, kf, vf, gf = (_split_heads(t.float().contiguous(), heads) for t in (q, k, v, d_out))   # [B,H,N,d]
        dq, dk, dv = torch.empty_like(qf), torch.empty_like(kf), torch.empty_like(vf)
        for h in range(heads):   # one head at a time: [B, N, N] fp32 blocks
            p = torch.softmax(torch.matmul(qf[:, h], kf[:, h].transpose(-1, -2)) * scale, dim=-1)
            dv[:, h] = torch.matmul(p.transpose(-1, -2), gf[:, h])
            dp = torch.matmul(gf[:, h], vf[:, h].transpose(-1, -2))
            ds = p * (dp - (p * dp).sum(dim=-1, keepdim=True))
            dq[:, h] = torch.matmul(ds, kf[:, h]) * scale
            dk[:, h] = torch.matmul(ds.transpose(-1, -2), qf[:, h]) * scale
        return _merge_heads(dq).to(q.dtype), _merge_heads(dk).to(k.dtype), _merge_heads(dv).to(v.dtype)


class CrossAttentionHeatFn(torch.autograd.Function):
    """(out, maps) = cross-attention + `_unravel_attn` (hook.py:28-56, 104-115): maps [B - b_first, T, N] fp32 is the
    mean over heads of the probabilities of the key tokens `token_idx` (all M when None)."""

    @staticmethod
    def forward(ctx, q, k, v, heads: int, scale: float, token_idx: Optional[Sequence[int]], b_first: int):
        B, N, _ = q.shape
        M = k.shape[1]
        T = M if token_idx is None else len(token_idx)
        maps = torch.empty((B - b_first, T, N), dtype=torch.float32, device=q.device)
        out = ops.attn_cross_heat(q, k, v, heads, maps, token_idx, b_first, accumulate=False, scale=scale)
        ctx.save_for_backward(q, k, v)
        ctx.heads, ctx.scale, ctx.b_first = heads, scale, b_first
        ctx.token_idx = None if token_idx is None else [int(t) for t in token_idx]
        return out, maps

    @staticmethod
    def backward(ctx, d_out, d_maps):
        q, k, v = ctx.saved_tensors
        if os.environ.get("AGENDA_TORCH_CROSS_BWD", "0") != "1":
            dq, dk, dv = ops.attn_cross_bwd(q, k, v, d_out.to(q.dtype), d_maps, ctx.heads, ctx.token_idx, ctx.b_first,
                                            scale=ctx.scale)
            return dq, dk, dv, None, None, None, None
        return CrossAttentionHeatFn._backward_torch(ctx, d_out, d_maps)

    @staticmethod
    def _backward_torch(ctx, d_out, d_maps):
        """The same gradients with cuBLAS batched GEMMs + softmax (kept as a cross-check: AGENDA_TORCH_CROSS_BWD=1)."""
        q, k, v = ctx.saved_tensors
        H = ctx.heads
        qf, kf, vf = (_split_heads(t.float(), H) for t in (q, k, v))           # [B,H,N|M,d]
        p = torch.softmax(torch.matmul(qf, kf.transpose(-1, -2)) * ctx.scale, dim=-1)   # [B,H,N,M]
        do = _split_heads(d_out.float().contiguous(), H)
        dv = torch.matmul(p.transpose(-1, -2), do)
        dp = torch.matmul(do, vf.transpose(-1, -2))
        if d_maps is not None:
            g = (d_maps.float() / H).permute(0, 2, 1).unsqueeze(1)              # [B',1,N,T]
            g = g.expand(-1, H, -1, -1)
            if ctx.token_idx is None:
                dp[ctx.b_first:] += g
            else:
                idx = torch.tensor(ctx.token_idx, device=dp.device, dtype=torch.long)
                dp[ctx.b_first:].index_add_(3, idx, g.contiguous())             # repeated tokens add up
        ds = p * (dp - (p * dp).sum(dim=-1, keepdim=True))
        dq = torch.matmul(ds, kf) * ctx.scale
        dk = torch.matmul(ds.transpose(-1, -2), qf) * ctx.scale
        return (_merge_heads(dq).to(q.dtype), _merge_heads(dk).to(k.dtype), _merge_heads(dv).to(v.dtype),
                None, None, None, None)


def global_heat_map_autograd(maps, latent_hw: int) -> torch.Tensor:
    """hook.py:59-81 on tensors that carry a graph: bicubic -> clamp(min=0) per map, stack, mean — torch ops, so the
    aggregation is differentiable exactly like the reference's."""
    if len(maps) == 0:
        raise RuntimeError('No heat maps found.')
    up = [F.interpolate(m.float(), size=(latent_hw, latent_hw), mode='bicubic').clamp_(min=0) for m in maps]
    return torch.stack(up, dim=0).mean(dim=0)
