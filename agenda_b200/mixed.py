"""Mixed-precision preparation of an fp32 checkpoint for the bf16 tensor-core path.

The reference loads its pipeline in fp32 (data_generation.py:30-31).  Running the UNet in bf16 rounds every weight
to 8 mantissa bits, which is harmless for the attention OUTPUTS (1e-2 tolerance) but not for the cross-attention
LOGITS the heat maps are made of: a bf16-rounded W_q / W_k alone moves the probabilities by up to ~2.5e-4.

`compensate_cross_projections(module)` is called on the fp32 model BEFORE `.to(torch.bfloat16)`: for the to_q and
to_k projections of every cross-attention module it registers the bf16 residual of the fp32 weight,

    weight_lo = bf16(W - bf16(W))            (W == bf16(W) + weight_lo to 2^-17 relative)

as a buffer next to the weight.  After the cast, `weight` holds bf16(W) and `weight_lo` the residual; the processor
(agenda_b200/processor.py, cross_logits="fp32") then evaluates to_q as two tensor-core GEMMs with fp32 accumulation
and fp32 output and the (once-per-prompt) key projection in fp32 from weight + weight_lo, so the logits that reach the
split-precision attention kernel are those of the fp32 checkpoint.  Modules without `weight_lo` work unchanged (their
bf16 weights are then taken as the exact weights, e.g. a checkpoint that was stored in bf16).
"""
from __future__ import annotations

import torch
from torch import nn


def _is_cross_attention(m: nn.Module) -> bool:
    q, k = getattr(m, "to_q", None), getattr(m, "to_k", None)
    if not isinstance(q, nn.Linear) or not isinstance(k, nn.Linear) or not hasattr(m, "heads"):
        return False
    if getattr(m, "is_cross", None) is not None:
        return bool(m.is_cross)
    if getattr(m, "is_cross_attention", None) is not None:   # diffusers Attention
        return bool(m.is_cross_attention)
    return q.in_features != k.in_features


def compensate_cross_projections(model: nn.Module) -> int:
    """Register `weight_lo` on to_q / to_k of every cross-attention module of `model` (fp32 weights expected).
    Returns the number of modules prepared.  Idempotent."""
    n = 0
    for m in model.modules():
        if not _is_cross_attention(m):
            continue
        for lin in (m.to_q, m.to_k):
            w = lin.weight.detach()
            if w.dtype != torch.float32:
                raise TypeError("compensate_cross_projections expects fp32 weights (call it before .to(bfloat16))")
            lo = (w - w.to(torch.bfloat16).float()).to(torch.bfloat16)
            if "weight_lo" in lin._buffers:
                lin.weight_lo = lo
            else:
                lin.register_buffer("weight_lo", lo)
        n += 1
    return n


def to_mixed_precision(model: nn.Module, dtype: torch.dtype = torch.bfloat16) -> nn.Module:
    """fp32 model -> `dtype` model whose cross-attention to_q / to_k keep their fp32 value as weight + weight_lo."""
    compensate_cross_projections(model)
    return model.to(dtype)
