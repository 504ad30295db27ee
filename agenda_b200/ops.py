"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, compute in libagenda_b200.so.

PyTorch is used for device memory and streams only.  Every function enqueues on torch's current stream and
raises if handed a CPU tensor (there is no CPU path in the product).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    """The C ABI launches on the calling thread's CURRENT device and torch's current stream of that device.  A caller
    whose tensors live on another GPU of the process (HeatmapPipeline(device="cuda:1") while cuda:0 is current, a
    device_map-sharded UNet) gets the right device made current for the duration of the call."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
            if isinstance(a, ContextKV):
                continue
        return fn(*args, **kwargs)
    return wrapper


def _dev(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"agenda_b200: `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"agenda_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"agenda_b200: unsupported attention dtype {t.dtype} (float32 or bfloat16)")


# ------------------------------------------------------------------ attention (hook.py:104-115) ---------------

@_on_tensor_device
def attn_self(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              precision: str = "bf16") -> torch.Tensor:
    """softmax(scale * q k^T) v per head.  q/k/v [B,N,H*d] token-major (to_q/to_k/to_v outputs).

    precision="bf16": tcgen05 tensor-core kernel (inputs are cast to bf16 if needed, fp32 softmax/accumulate).
    precision="fp32": exact fp32 CUDA-core kernel.
    """
    q = _dev(q, "q")
    B, N, C = q.shape
    d = C // heads
    scale = float(d ** -0.5 if scale is None else scale)
    if precision == "bf16" and d not in (40, 64, 80, 160):
        # head dims the tensor-core kernels are not instantiated for: the exact fp32 kernel (any d <= 160), same contract
        odt = q.dtype
        if odt not in (torch.float32, torch.bfloat16):
            q, k, v = (t.to(torch.bfloat16) for t in (q, k, v))
        out = attn_self(q, k, v, heads, scale, precision="fp32")
        return out if out.dtype == odt else out.to(odt)
    if precision == "bf16":
        odt = q.dtype
        qb, kb, vb = (_dev(t, n).to(torch.bfloat16) for t, n in ((q, "q"), (k, "k"), (v, "v")))
        out = torch.empty_like(qb)
        _lib.call("agenda_attn_self_fwd", qb.data_ptr(), kb.data_ptr(), vb.data_ptr(), out.data_ptr(), _lib.BF16,
                  B, heads, N, d, scale, _stream())
        return out if odt == torch.bfloat16 else out.to(odt)
    if precision != "fp32":
        raise ValueError("precision must be 'bf16' or 'fp32'")
    k, v = _dev(k, "k", q.dtype), _dev(v, "v", q.dtype)
    out = torch.empty_like(q)
    _lib.call("agenda_attn_self_fwd_f32", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _dtype_code(q),
              B, heads, N, d, scale, _stream())
    return out


@_on_tensor_device
def attn_self_fused_qkv(qkv: torch.Tensor, heads: int, scale: Optional[float] = None,
                        prescaled: bool = False) -> torch.Tensor:
    """Self-attention on the output of ONE fused q/k/v projection: qkv bf16 [B,N,3C] with q = [..,:C], k = [..,C:2C],
    v = [..,2C:].  The kernel reads the three column slices in place (row stride 3C); returns out [B,N,C]."""
    qkv = _dev(qkv, "qkv", torch.bfloat16)
    B, N, C3 = qkv.shape
    if C3 % 3:
        raise ValueError("qkv last dim must be 3*C")
    C = C3 // 3
    d = C // heads
    # prescaled: the q columns already carry scale * log2(e) (folded into W_q) -> the ABI's scale == 0 convention
    scale = 0.0 if prescaled else float(d ** -0.5 if scale is None else scale)
    out = torch.empty((B, N, C), dtype=torch.bfloat16, device=qkv.device)
    p = qkv.data_ptr()
    _lib.call("agenda_attn_self_fwd_strided", p, p + 2 * C, p + 4 * C, out.data_ptr(), _lib.BF16, B, heads, N, d, C3,
              scale, _stream())
    return out


@_on_tensor_device
def attn_masked(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, mask: torch.Tensor,
                maps: Optional[torch.Tensor] = None, token_idx: Optional[Sequence[int]] = None, b_first: int = 0,
                accumulate: bool = False, scale: Optional[float] = None, per_head: bool = False) -> torch.Tensor:
    """softmax(scale q k^T + mask) v with the additive mask `attn.prepare_attention_mask` returns (hook.py:92,108):
    mask [B*heads, 1 | N, M] (any float dtype; bool / 0-1 masks must already be additive, as diffusers makes them).
    Self- and cross-attention alike; `maps` etc. as in attn_cross_heat.  Exact fp32 CUDA-core path."""
    q = _dev(q, "q")
    if q.dtype == torch.float16:
        return attn_masked(q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16), heads, mask, maps, token_idx,
                           b_first, accumulate, scale, per_head).to(torch.float16)
    k, v = _dev(k, "k", q.dtype), _dev(v, "v", q.dtype)
    B, N, C = q.shape
    M = k.shape[1]
    d = C // heads
    scale = float(d ** -0.5 if scale is None else scale)
    if mask.dim() != 3 or mask.shape[0] != B * heads or mask.shape[1] not in (1, N) or mask.shape[2] != M:
        raise ValueError(f"mask must be [B*heads={B * heads}, 1 or N={N}, M={M}], got {tuple(mask.shape)}")
    mask = mask.to(device=q.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(q)
    if maps is not None:
        maps = _dev(maps, "maps", torch.float32)
        T = M if token_idx is None else len(token_idx)
        lead = (B - b_first, heads, T) if per_head else (B - b_first, T)
        if tuple(maps.shape[:len(lead)]) != lead or maps.numel() != N * int(torch.tensor(lead).prod()):
            raise ValueError(f"maps must be {list(lead)} + [{N}], got {tuple(maps.shape)}")
        if not maps.is_contiguous():
            raise ValueError("maps must be contiguous (it is written in place)")
        idx = None if token_idx is None else (ctypes.c_int32 * T)(*[int(i) for i in token_idx])
        mp = maps.data_ptr()
    else:
        T, idx, mp = 0, None, None
    _lib.call("agenda_attn_fwd_masked", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _dtype_code(q),
              B, heads, N, M, d, scale, mask.data_ptr(), int(mask.shape[1]), idx, T, int(b_first), int(bool(per_head)), mp,
              int(bool(accumulate)), _stream())
    return out


@_on_tensor_device
def attn_cross_heat(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, maps: Optional[torch.Tensor],
                    token_idx: Optional[Sequence[int]] = None, b_first: int = 0, accumulate: bool = False,
                    scale: Optional[float] = None, force_f32_kernel: bool = False,
                    per_head: bool = False) -> torch.Tensor:
    """Cross-attention + heat epilogue (hook.py:108-114 and _unravel_attn hook.py:28-56).

    maps: fp32 [B-b_first, T, N] written (accumulate=False) or added to (accumulate=True), T = len(token_idx) or
    M when token_idx is None; pass maps=None to skip the epilogue.  per_head=True keeps the heads apart (DAAM):
    maps is [B-b_first, H, T, N] and no head mean is taken.  Returns out [B,N,H*d]."""
    q = _dev(q, "q")
    k, v = _dev(k, "k", q.dtype), _dev(v, "v", q.dtype)
    if q.dtype == torch.float16:
        # fp16 pipelines (torch_dtype=torch.float16): the tensor-core kernels take bf16 (same 16-bit storage, fp32
        # softmax and accumulation inside); the attention output goes back to fp16, the heat maps are fp32 anyway
        out = attn_cross_heat(q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16), heads, maps, token_idx,
                              b_first, accumulate, scale, force_f32_kernel, per_head)
        return out.to(torch.float16)
    B, N, C = q.shape
    M = k.shape[1]
    d = C // heads
    scale = float(d ** -0.5 if scale is None else scale)
    out = torch.empty_like(q)
    if maps is not None:
        maps = _dev(maps, "maps", torch.float32)
        T = M if token_idx is None else len(token_idx)
        lead = (B - b_first, heads, T) if per_head else (B - b_first, T)
        if tuple(maps.shape[:len(lead)]) != lead or maps.numel() != N * int(torch.tensor(lead).prod()):
            raise ValueError(f"maps must be {list(lead)} + [{N}] (any trailing shape of {N} elements), "
                             f"got {tuple(maps.shape)}")
        if not maps.is_contiguous():
            raise ValueError("maps must be contiguous (it is written in place)")
        idx = None if token_idx is None else (ctypes.c_int32 * T)(*[int(i) for i in token_idx])
        mp = maps.data_ptr()
    else:
        T, idx, mp = 0, None, None
    if per_head:
        if maps is None or force_f32_kernel:
            raise ValueError("per_head=True needs maps and has no forced-fp32 variant")
        entry = "agenda_attn_cross_fwd_heat_heads"
    else:
        entry = "agenda_attn_cross_fwd_heat_f32" if force_f32_kernel else "agenda_attn_cross_fwd_heat"
    _lib.call(entry, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _dtype_code(q),
              B, heads, N, M, d, scale, idx, T, int(b_first), mp, int(bool(accumulate)), _stream())
    return out


SELF_BWD_HEAD_DIMS = (40, 64, 80, 160)


@_on_tensor_device
def attn_self_with_lse(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None):
    """attn_self on the tensor cores that also returns the rows' base-2 log-sum-exp [B,H,N] fp32 for attn_self_bwd (training
    mode), or (out, None) where the kernel that runs does not emit it (N <= 128, other head dims)."""
    q = _dev(q, "q")
    B, N, C = q.shape
    d = C // heads
    if d not in (40, 64, 80, 160) or not _lib.load().agenda_attn_self_fwd_emits_lse(N, d):
        return attn_self(q, k, v, heads, scale), None
    scale = float(d ** -0.5 if scale is None else scale)
    odt = q.dtype
    qb, kb, vb = (_dev(t, n).to(torch.bfloat16) for t, n in ((q, "q"), (k, "k"), (v, "v")))
    out = torch.empty_like(qb)
    lse = torch.empty((B, heads, N), dtype=torch.float32, device=q.device)
    _lib.call("agenda_attn_self_fwd_lse", qb.data_ptr(), kb.data_ptr(), vb.data_ptr(), out.data_ptr(), lse.data_ptr(), _lib.BF16,
              B, heads, N, d, scale, _stream())
    return (out if odt == torch.bfloat16 else out.to(odt)), lse


@_on_tensor_device
def attn_self_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor, heads: int,
                  scale: Optional[float] = None, lse: Optional[torch.Tensor] = None):
    """Backward of attn_self on the tensor cores (agenda_attn_self_bwd): q/k/v/out/d_out [B,N,H*d] -> (dq, dk, dv) in
    the dtype of q.  Inputs are taken as bf16 (cast if needed); P is recomputed from a log-sum-exp pass."""
    q = _dev(q, "q")
    odt = q.dtype
    qb, kb, vb, ob, gb = (_dev(t, n).to(torch.bfloat16).contiguous()
                          for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (d_out, "d_out")))
    B, N, C = qb.shape
    d = C // heads
    scale = float(d ** -0.5 if scale is None else scale)
    ws = torch.empty(_lib.load().agenda_attn_self_bwd_workspace_bytes(B, heads, N) // 4, dtype=torch.float32, device=q.device)
    dq, dk, dv = torch.empty_like(qb), torch.empty_like(qb), torch.empty_like(qb)
    if lse is not None:   # from attn_self_with_lse: the backward skips its own log-sum-exp pass
        lse = _dev(lse, "lse", torch.float32)
        if tuple(lse.shape) != (B, heads, N):
            raise ValueError(f"lse must be [{B}, {heads}, {N}]")
        _lib.call("agenda_attn_self_bwd_lse", qb.data_ptr(), kb.data_ptr(), vb.data_ptr(), ob.data_ptr(), gb.data_ptr(),
                  lse.data_ptr(), ws.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), _lib.BF16, B, heads, N, d, scale,
                  _stream())
    else:
        _lib.call("agenda_attn_self_bwd", qb.data_ptr(), kb.data_ptr(), vb.data_ptr(), ob.data_ptr(), gb.data_ptr(),
                  ws.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), _lib.BF16, B, heads, N, d, scale, _stream())
    if odt != torch.bfloat16:
        dq, dk, dv = dq.to(odt), dk.to(odt), dv.to(odt)
    return dq, dk, dv


def split_bf16(x: torch.Tensor):
    """x (fp32) -> (hi, lo) bf16 with hi = bf16(x), lo = bf16(x - hi): x == hi + lo to 2^-17 relative (the operand form
    of the split-precision cross-attention kernel; pack_context_kv does the same on the device)."""
    x = x.float()
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def linear_split_f32_supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.dim() == 2
            and w.shape[1] % 64 == 0 and w.shape[0] % 160 == 0 and x.shape[-1] == w.shape[1])


@_on_tensor_device
def linear_split_f32(x: torch.Tensor, w_hi: torch.Tensor, w_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 [..., K] @ (w_hi + w_lo)[N, K]^T -> fp32 [..., N] in one tcgen05 GEMM (agenda_linear_split_f32): fp32
    accumulation AND fp32 output, both weight halves into the same accumulator.  w_lo None: plain fp32-output GEMM."""
    x = _dev(x, "x", torch.bfloat16)
    w_hi = _dev(w_hi, "w_hi", torch.bfloat16)
    if w_lo is not None:
        w_lo = _dev(w_lo, "w_lo", torch.bfloat16)
        if w_lo.shape != w_hi.shape:
            raise ValueError("w_lo must have the shape of w_hi")
    N, K = w_hi.shape
    if x.shape[-1] != K:
        raise ValueError(f"x [..., {x.shape[-1]}] does not match the weight [{N}, {K}]")
    M = x.numel() // K
    out = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
    _lib.call("agenda_linear_split_f32", x.data_ptr(), w_hi.data_ptr(), None if w_lo is None else w_lo.data_ptr(),
              out.data_ptr(), M, K, N, _stream())
    return out


class PackedSplitWeight:
    """w_hi | w_lo of a linear layer pre-packed for agenda_linear_split_f32_packed (agenda_linear_split_pack_w)."""

    def __init__(self, blob: torch.Tensor, N: int, K: int, has_lo: bool):
        self.blob, self.N, self.K, self.has_lo = blob, N, K, has_lo


@_on_tensor_device
def linear_split_pack(w_hi: torch.Tensor, w_lo: Optional[torch.Tensor] = None) -> PackedSplitWeight:
    """Lay the bf16 weight halves [N, K] out as the GEMM's pipeline-stage images (one bulk copy per K block)."""
    w_hi = _dev(w_hi, "w_hi", torch.bfloat16)
    if w_lo is not None:
        w_lo = _dev(w_lo, "w_lo", torch.bfloat16)
        if w_lo.shape != w_hi.shape:
            raise ValueError("w_lo must have the shape of w_hi")
    N, K = w_hi.shape
    nbytes = _lib.load().agenda_linear_split_pack_bytes(N, K, int(w_lo is not None))
    if nbytes < 0:
        raise _lib.AgendaError(int(nbytes), _lib.load().agenda_last_error().decode("utf-8", "replace"))
    blob = torch.empty(nbytes, dtype=torch.uint8, device=w_hi.device)
    _lib.call("agenda_linear_split_pack_w", w_hi.data_ptr(), None if w_lo is None else w_lo.data_ptr(), blob.data_ptr(), N, K,
              _stream())
    return PackedSplitWeight(blob, N, K, w_lo is not None)


@_on_tensor_device
def linear_split_f32_packed(x: torch.Tensor, w: PackedSplitWeight) -> torch.Tensor:
    """linear_split_f32 with pre-packed weights: same numbers, the weights of a K block arrive with one bulk copy."""
    x = _dev(x, "x", torch.bfloat16)
    if x.shape[-1] != w.K:
        raise ValueError(f"x [..., {x.shape[-1]}] does not match the packed weight [{w.N}, {w.K}]")
    M = x.numel() // w.K
    out = torch.empty(x.shape[:-1] + (w.N,), dtype=torch.float32, device=x.device)
    _lib.call("agenda_linear_split_f32_packed", x.data_ptr(), w.blob.data_ptr(), int(w.has_lo), out.data_ptr(), M, w.K, w.N,
              _stream())
    return out


class QueryChunks:
    """An fp32 query projection in the chunk-major layout [B][H][d/40][N][40] (agenda_linear_split_f32_heads), the form
    agenda_attn_cross_fwd_heat_x3_hm streams with one bulk copy per 128-query chunk.  `to_rows()` gives [B,N,H*d] back."""

    def __init__(self, buf: torch.Tensor, B: int, N: int, heads: int, d: int):
        self.buf, self.B, self.N, self.heads, self.d = buf, B, N, heads, d

    @property
    def shape(self):
        return (self.B, self.N, self.heads * self.d)

    def to_rows(self) -> torch.Tensor:
        B, N, H, d = self.B, self.N, self.heads, self.d
        return self.buf.view(B, H, d // 40, N, 40).permute(0, 3, 1, 2, 4).reshape(B, N, H * d)


@_on_tensor_device
def linear_split_f32_heads(x: torch.Tensor, w_hi: torch.Tensor, w_lo: Optional[torch.Tensor], heads: int) -> QueryChunks:
    """linear_split_f32 for x bf16 [B, N, K] with the fp32 result in the chunk-major layout of the cross-attention kernel
    (agenda_linear_split_f32_heads); head dim = w.shape[0] / heads must be a multiple of 40."""
    x = _dev(x, "x", torch.bfloat16)
    w_hi = _dev(w_hi, "w_hi", torch.bfloat16)
    if w_lo is not None:
        w_lo = _dev(w_lo, "w_lo", torch.bfloat16)
        if w_lo.shape != w_hi.shape:
            raise ValueError("w_lo must have the shape of w_hi")
    if x.dim() != 3:
        raise ValueError("x must be [B, N, K]")
    B, Nq, K = x.shape
    Nout = w_hi.shape[0]
    if w_hi.shape[1] != K or Nout % heads or (Nout // heads) % 40:
        raise ValueError(f"weight {tuple(w_hi.shape)} does not give {heads} heads of a multiple of 40 columns from K={K}")
    buf = torch.empty(B * Nq * Nout, dtype=torch.float32, device=x.device)
    _lib.call("agenda_linear_split_f32_heads", x.data_ptr(), w_hi.data_ptr(), None if w_lo is None else w_lo.data_ptr(),
              buf.data_ptr(), B * Nq, K, Nout, Nq, heads, _stream())
    return QueryChunks(buf, B, Nq, heads, Nout // heads)


class ContextKV:
    """The prompt side of the split-precision cross-attention, packed for the tensor cores (agenda_pack_context_kv):
    `blob` u8 [B, H, block] with K_hi | K_lo | V of every (batch, head) in the kernel's shared-memory layout."""

    def __init__(self, blob: torch.Tensor, B: int, M: int, heads: int, d: int):
        self.blob, self.B, self.M, self.heads, self.d = blob, B, M, heads, d


@_on_tensor_device
def pack_context_kv(k32: torch.Tensor, v: torch.Tensor, heads: int, out: Optional[ContextKV] = None) -> ContextKV:
    """k32 fp32 [B,M,H*d] (the key projection in fp32), v fp32 or bf16 [B,M,H*d] -> ContextKV.  Pass `out` to refill an
    existing blob in place (same shapes): captured CUDA graphs keep reading the same buffer."""
    k32 = _dev(k32, "k32", torch.float32)
    v = _dev(v, "v")
    if v.dtype not in (torch.float32, torch.bfloat16):
        v = v.to(torch.bfloat16)
    B, M, C = k32.shape
    if tuple(v.shape) != (B, M, C) or C % heads:
        raise ValueError(f"k32 / v must both be [B,M,H*d], got {tuple(k32.shape)} {tuple(v.shape)}, heads={heads}")
    d = C // heads
    nbytes = _lib.load().agenda_context_blob_bytes(B, heads, d)
    if nbytes < 0:
        raise _lib.AgendaError(int(nbytes), _lib.load().agenda_last_error().decode("utf-8", "replace"))
    if out is None:
        out = ContextKV(torch.empty((B, heads, nbytes // (B * heads)), dtype=torch.uint8, device=k32.device), B, M, heads, d)
    elif (out.B, out.M, out.heads, out.d) != (B, M, heads, d) or out.blob.numel() != nbytes:
        raise ValueError("pack_context_kv: `out` was built for a different shape")
    _lib.call("agenda_pack_context_kv", k32.data_ptr(), v.data_ptr(), _dtype_code(v), out.blob.data_ptr(), B, heads, M, d,
              _stream())
    return out


@_on_tensor_device
def attn_cross_heat_x3(q: torch.Tensor, ctx: ContextKV, maps: Optional[torch.Tensor],
                       token_idx: Optional[Sequence[int]] = None, b_first: int = 0, accumulate: bool = False,
                       scale: Optional[float] = None, per_head: bool = False,
                       out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """Cross-attention + heat epilogue with fp32-accurate logits on the bf16 tensor cores (hook.py:108-114 at the
    reference's fp32 precision; heat maps to ~1e-6 of an fp32 softmax).

    q fp32 [B,N,H*d] (to_q output with fp32 accumulation); ctx = pack_context_kv(K fp32, V).  maps / token_idx / b_first
    / accumulate / per_head as in attn_cross_heat (token_idx None = all prompt tokens; per_head takes at most 8).
    Returns out [B,N,H*d] in `out_dtype` (bf16 or fp32)."""
    chunked = isinstance(q, QueryChunks)
    if chunked:
        if q.heads != ctx.heads or q.d != ctx.d:
            raise ValueError("the chunk-major query was projected for another head layout than the packed context")
        qbuf = q.buf
    else:
        q = _dev(q, "q", torch.float32)
        qbuf = q
    B, N, C = q.shape
    heads, M, d = ctx.heads, ctx.M, ctx.d
    if ctx.B != B or heads * d != C or not ctx.blob.is_cuda:
        raise ValueError(f"context was packed for B={ctx.B}, H*d={heads * d}; q is {tuple(q.shape)}")
    scale = float(d ** -0.5 if scale is None else scale)
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out_dtype must be bfloat16 or float32")
    out = torch.empty((B, N, C), dtype=out_dtype, device=qbuf.device)
    if maps is not None:
        maps = _dev(maps, "maps", torch.float32)
        T = M if token_idx is None else len(token_idx)
        lead = (B - b_first, heads, T) if per_head else (B - b_first, T)
        n_lead = 1
        for x in lead:
            n_lead *= x
        if tuple(maps.shape[:len(lead)]) != lead or maps.numel() != N * n_lead:
            raise ValueError(f"maps must be {list(lead)} + [{N}] (any trailing shape of {N} elements), "
                             f"got {tuple(maps.shape)}")
        if not maps.is_contiguous():
            raise ValueError("maps must be contiguous (it is written in place)")
        idx = None if token_idx is None else (ctypes.c_int32 * T)(*[int(i) for i in token_idx])
        mp = maps.data_ptr()
    else:
        T, idx, mp = 0, None, None
    _lib.call("agenda_attn_cross_fwd_heat_x3_hm" if chunked else "agenda_attn_cross_fwd_heat_x3", qbuf.data_ptr(),
              ctx.blob.data_ptr(), out.data_ptr(),
              _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32, B, heads, N, M, d, scale, idx, T,
              int(b_first), int(bool(per_head)), mp, int(bool(accumulate)), _stream())
    return out


@_on_tensor_device
def attn_cross_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, d_out: torch.Tensor,
                   d_maps: Optional[torch.Tensor], heads: int, token_idx: Optional[Sequence[int]] = None,
                   b_first: int = 0, scale: Optional[float] = None):
    """Backward of attn_cross_heat (training mode): returns (dq, dk, dv) in the dtypes of q / k / v.  d_maps is the
    gradient of the head-mean maps, fp32 [B-b_first, T, N] (or any trailing shape of N elements), or None."""
    q = _dev(q, "q")
    if q.dtype == torch.float16:
        # fp16 training (accelerate mixed_precision=fp16, finetune_sd_token.py): the kernel computes in fp32 anyway, so the
        # fp16 tensors go in as fp32 (exact) and the gradients come back in fp16, like the forward pass handles fp16
        dq, dk, dv = attn_cross_bwd(q.float(), k.float(), v.float(), d_out.float(), d_maps, heads, token_idx, b_first, scale)
        return dq.to(torch.float16), dk.to(torch.float16), dv.to(torch.float16)
    k, v, d_out = _dev(k, "k", q.dtype), _dev(v, "v", q.dtype), _dev(d_out, "d_out", q.dtype)
    B, N, C = q.shape
    M = k.shape[1]
    d = C // heads
    scale = float(d ** -0.5 if scale is None else scale)
    dq = torch.empty_like(q)
    dk = torch.zeros((B, M, C), dtype=torch.float32, device=q.device)
    dv = torch.zeros_like(dk)
    if d_maps is not None:
        d_maps = _dev(d_maps, "d_maps", torch.float32)
        T = M if token_idx is None else len(token_idx)
        if d_maps.numel() != (B - b_first) * T * N:
            raise ValueError(f"d_maps must hold {(B - b_first)}x{T}x{N} elements, got {tuple(d_maps.shape)}")
        idx = None if token_idx is None else (ctypes.c_int32 * T)(*[int(i) for i in token_idx])
        mp = d_maps.data_ptr()
    else:
        T, idx, mp = 0, None, None
    # bf16 tensors with at most 8 selected tokens (every caller of the reference) take the tensor-core kernels; fp32
    # tensors, all-token maps and AGENDA_CROSS_BWD_TC=0 the exact fp32 CUDA-core kernel
    use_tc = (q.dtype == torch.bfloat16 and M <= 80 and d in (40, 64, 80, 160) and (mp is None or 1 <= T <= 8)
              and os.environ.get("AGENDA_CROSS_BWD_TC", "1") != "0")
    if use_tc:
        ws = torch.empty(_lib.load().agenda_attn_cross_bwd_tc_workspace_bytes(B, heads, N) // 4, dtype=torch.float32,
                         device=q.device)
        _lib.call("agenda_attn_cross_bwd_tc", q.data_ptr(), k.data_ptr(), v.data_ptr(), d_out.data_ptr(), mp, ws.data_ptr(),
                  dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, heads, N, M, d, scale, idx, T, int(b_first), _stream())
    else:
        _lib.call("agenda_attn_cross_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), d_out.data_ptr(), mp, dq.data_ptr(),
                  dk.data_ptr(), dv.data_ptr(), _dtype_code(q), B, heads, N, M, d, scale, idx, T, int(b_first), _stream())
    return dq, dk.to(k.dtype), dv.to(v.dtype)


# ------------------------------------------------------------------ heat maps (hook.py:59-81) -----------------

@_on_tensor_device
def heat_upsample_accum(maps: torch.Tensor, acc: torch.Tensor) -> None:
    """acc[..., L, L] += clamp(bicubic(maps[..., h, w] -> L x L), min=0)  (hook.py:72), in place."""
    maps, acc = _dev(maps, "maps", torch.float32), _dev(acc, "acc", torch.float32)
    if not acc.is_contiguous():
        raise ValueError("acc must be contiguous")
    h, w = maps.shape[-2:]
    L = acc.shape[-1]
    n = maps.numel() // (h * w)
    if acc.shape[-2] != L or acc.numel() != n * L * L:
        raise ValueError(f"acc {tuple(acc.shape)} does not match maps {tuple(maps.shape)}")
    _lib.call("agenda_heat_upsample_accum", maps.data_ptr(), acc.data_ptr(), n, h, w, L, _stream())


@_on_tensor_device
def heat_upsample_accum_heads(maps: torch.Tensor, acc: torch.Tensor) -> None:
    """DAAM-style aggregation: maps fp32 [B', G, T, h, w] (G = heads) -> acc fp32 [B', T, L, L] +=
    sum_g clamp(bicubic(maps[:, g]), min=0), every (batch, head, token) plane upsampled and clamped on its own."""
    maps, acc = _dev(maps, "maps", torch.float32), _dev(acc, "acc", torch.float32)
    if maps.dim() != 5 or acc.dim() != 4 or not acc.is_contiguous():
        raise ValueError("maps must be [B',G,T,h,w] and acc a contiguous [B',T,L,L]")
    Bp, G, T, h, w = maps.shape
    L = acc.shape[-1]
    if tuple(acc.shape) != (Bp, T, L, L):
        raise ValueError(f"acc {tuple(acc.shape)} does not match maps {tuple(maps.shape)}")
    _lib.call("agenda_heat_upsample_accum_heads", maps.data_ptr(), acc.data_ptr(), Bp * T, T, G, h, w, L, _stream())


@_on_tensor_device
def heat_finalize(acc: torch.Tensor, count: int) -> torch.Tensor:
    """acc / count (the mean over the (layer x step) list, hook.py:79)."""
    acc = _dev(acc, "acc", torch.float32)
    out = torch.empty_like(acc)
    _lib.call("agenda_heat_finalize", acc.data_ptr(), out.data_ptr(), acc.numel(), int(count), _stream())
    return out


# ------------------------------------------------------------------ post-processing ----------------------------

@_on_tensor_device
def heat_normalize_u8(heat: torch.Tensor) -> torch.Tensor:
    """data_generation.py:82 + astype(uint8): heat fp32 [..., h, w] -> u8 same shape, per-map min-max."""
    heat = _dev(heat, "heat", torch.float32)
    h, w = heat.shape[-2:]
    out = torch.empty(heat.shape, dtype=torch.uint8, device=heat.device)
    _lib.call("agenda_heat_normalize_u8", heat.data_ptr(), out.data_ptr(), heat.numel() // (h * w), h * w, _stream())
    return out


@_on_tensor_device
def resize_bicubic_u8(img: torch.Tensor, size) -> torch.Tensor:
    """PIL Image.resize((W,H)) default filter for mode 'L' (data_generation.py:85), bit-exact.  img u8 [...,h,w]."""
    img = _dev(img, "img", torch.uint8)
    Ho, Wo = (size, size) if isinstance(size, int) else size
    h, w = img.shape[-2:]
    n = img.numel() // (h * w)
    out = torch.empty(img.shape[:-2] + (Ho, Wo), dtype=torch.uint8, device=img.device)
    _lib.call("agenda_resize_bicubic_u8", img.data_ptr(), out.data_ptr(), n, h, w, Ho, Wo, _stream())
    return out


@_on_tensor_device
def heat_to_u8_image(heat: torch.Tensor, size) -> torch.Tensor:
    """data_generation.py:82-85 fused: fp32 [...,h,w] -> min-max -> u8 -> PIL-bicubic resize -> u8 [...,Ho,Wo]."""
    heat = _dev(heat, "heat", torch.float32)
    Ho, Wo = (size, size) if isinstance(size, int) else size
    h, w = heat.shape[-2:]
    n = heat.numel() // (h * w)
    out = torch.empty(heat.shape[:-2] + (Ho, Wo), dtype=torch.uint8, device=heat.device)
    _lib.call("agenda_heat_to_u8_image", heat.data_ptr(), out.data_ptr(), n, h, w, Ho, Wo, _stream())
    return out


@_on_tensor_device
def stack_heatmaps_u8(obj: torch.Tensor, fg: torch.Tensor, bg: torch.Tensor):
    """postprocess_heatmap.py:44-46: returns (stack u8 [...,H,W,3], inv_bg u8 [...,H,W])."""
    obj, fg, bg = (_dev(t, n, torch.uint8) for t, n in ((obj, "obj"), (fg, "fg"), (bg, "bg")))
    if not (obj.shape == fg.shape == bg.shape):
        raise ValueError("obj/fg/bg shapes differ")
    H, W = obj.shape[-2:]
    n = obj.numel() // (H * W)
    stack = torch.empty(obj.shape + (3,), dtype=torch.uint8, device=obj.device)
    inv = torch.empty_like(obj)
    _lib.call("agenda_stack_heatmaps_u8", obj.data_ptr(), fg.data_ptr(), bg.data_ptr(), stack.data_ptr(),
              inv.data_ptr(), n, H, W, _stream())
    return stack, inv


@_on_tensor_device
def heat_postprocess_stack(heat: torch.Tensor, size):
    """a7+a8 fused for (object, fg, bg) triples: heat fp32 [n,3,h,w] -> (planes u8 [n,3,Ho,Wo], stack u8
    [n,Ho,Wo,3], inv_bg u8 [n,Ho,Wo])."""
    heat = _dev(heat, "heat", torch.float32)
    if heat.dim() != 4 or heat.shape[1] != 3:
        raise ValueError("heat must be [n,3,h,w]")
    Ho, Wo = (size, size) if isinstance(size, int) else size
    n, _, h, w = heat.shape
    planes = torch.empty((n, 3, Ho, Wo), dtype=torch.uint8, device=heat.device)
    stack = torch.empty((n, Ho, Wo, 3), dtype=torch.uint8, device=heat.device)
    inv = torch.empty((n, Ho, Wo), dtype=torch.uint8, device=heat.device)
    _lib.call("agenda_heat_postprocess_stack", heat.data_ptr(), planes.data_ptr(), stack.data_ptr(), inv.data_ptr(),
              n, h, w, Ho, Wo, _stream())
    return planes, stack, inv


def groupnorm_nhwc_supported(x: torch.Tensor, groups: int) -> bool:
    """A bf16 CUDA tensor [B,C,H,W] in channels-last memory (or [B,HW,C] contiguous) the NHWC GroupNorm kernel takes."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() in (3, 4)):
        return False
    C = x.shape[1] if x.dim() == 4 else x.shape[2]
    if C % 8 or C % groups or groups > 64:
        return False
    return x.permute(0, 2, 3, 1).is_contiguous() if x.dim() == 4 else x.is_contiguous()


@_on_tensor_device
def groupnorm_nhwc(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], groups: int,
                   eps: float = 1e-5, silu: bool = False, pre_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.nn.GroupNorm (+ SiLU) on a channels-last bf16 tensor without leaving the channels-last layout
    (agenda_groupnorm_nhwc).  x [B,C,H,W] channels-last -> same shape / strides; x [B,HW,C] -> [B,HW,C].
    pre_add [B,C]: added to x (broadcast over the pixels) before the normalisation."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("agenda_b200: `x` must be a CUDA tensor (no CPU fallback exists)")
    if not groupnorm_nhwc_supported(x, groups):   # (no .contiguous() here: that would turn channels-last into NCHW)
        raise ValueError("groupnorm_nhwc: need a channels-last bf16 tensor with C % 8 == 0 and C % groups == 0")
    if x.dim() == 4:
        B, C, H, W = x.shape
        HW = H * W
        y = torch.empty_like(x)            # preserves the channels-last strides
    else:
        B, HW, C = x.shape
        y = torch.empty_like(x)
    ws = torch.empty(_lib.load().agenda_groupnorm_workspace_bytes(B, HW, C, groups) // 4, dtype=torch.float32, device=x.device)
    w = None if weight is None else _dev(weight, "weight", torch.bfloat16).contiguous()
    b = None if bias is None else _dev(bias, "bias", torch.bfloat16).contiguous()
    pa = None
    if pre_add is not None:
        pa = _dev(pre_add, "pre_add", torch.bfloat16)
        if tuple(pa.shape) != (B, C):
            raise ValueError("groupnorm_nhwc: pre_add must be [B, C]")
    _lib.call("agenda_groupnorm_nhwc", x.data_ptr(), 0 if w is None else w.data_ptr(), 0 if b is None else b.data_ptr(),
              0 if pa is None else pa.data_ptr(), y.data_ptr(), ws.data_ptr(), B, HW, C, groups, float(eps),
              1 if silu else 0, _stream())
    return y


@_on_tensor_device
def add_bias_residual(h: torch.Tensor, bias: Optional[torch.Tensor], res: torch.Tensor) -> torch.Tensor:
    """h + bias[c] + res for channels-last bf16 [B,C,H,W] tensors of identical strides (agenda_add_bias_residual)."""
    if not (h.is_cuda and h.dtype == torch.bfloat16 and res.dtype == torch.bfloat16 and h.shape == res.shape
            and h.stride() == res.stride() and h.dim() == 4 and h.permute(0, 2, 3, 1).is_contiguous()):
        raise ValueError("add_bias_residual: need two channels-last bf16 CUDA tensors of the same shape")
    B, C, H, W = h.shape
    y = torch.empty_like(h)
    b = None if bias is None else _dev(bias, "bias", torch.bfloat16)
    _lib.call("agenda_add_bias_residual", h.data_ptr(), 0 if b is None else b.data_ptr(), res.data_ptr(), y.data_ptr(),
              B * H * W, C, _stream())
    return y


@_on_tensor_device
def layernorm(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    """torch.nn.LayerNorm over the last dim of a bf16 tensor (agenda_layernorm): one warp per row, fp32 statistics."""
    x = _dev(x, "x", torch.bfloat16)
    C = x.shape[-1]
    if C % 8:
        raise ValueError("layernorm: last dim must be a multiple of 8")
    y = torch.empty_like(x)
    w = None if weight is None else _dev(weight, "weight", torch.bfloat16)
    b = None if bias is None else _dev(bias, "bias", torch.bfloat16)
    _lib.call("agenda_layernorm", x.data_ptr(), 0 if w is None else w.data_ptr(), 0 if b is None else b.data_ptr(),
              y.data_ptr(), x.numel() // C, C, float(eps), _stream())
    return y


@_on_tensor_device
def geglu(x: torch.Tensor) -> torch.Tensor:
    """x [..., 2*inner] = [a | gate] (bf16) -> a * gelu(gate) [..., inner] (agenda_geglu; diffusers GEGLU.forward after
    its projection)."""
    x = _dev(x, "x", torch.bfloat16).contiguous()
    inner = x.shape[-1] // 2
    if x.shape[-1] % 16:
        raise ValueError("geglu: last dim must be a multiple of 16")
    y = torch.empty(x.shape[:-1] + (inner,), dtype=x.dtype, device=x.device)
    _lib.call("agenda_geglu", x.data_ptr(), y.data_ptr(), x.numel() // x.shape[-1], inner, _stream())
    return y


@_on_tensor_device
def ccl_bbox(heat: torch.Tensor, thr: float = 0.5, max_boxes: int = 256, want_labels: bool = True):
    """Threshold + 4-connected components + boxes.  heat fp32 [n,H,W] ->
    (labels int32 [n,H,W] or None, counts int32 [n], boxes int32 [n,max_boxes,5] = x,y,w,h,area)."""
    heat = _dev(heat, "heat", torch.float32)
    if heat.dim() == 2:
        heat = heat[None]
    n, H, W = heat.shape
    labels = torch.empty((n, H, W), dtype=torch.int32, device=heat.device) if want_labels else None
    counts = torch.empty((n,), dtype=torch.int32, device=heat.device)
    boxes = torch.zeros((n, max_boxes, 5), dtype=torch.int32, device=heat.device)
    _lib.call("agenda_ccl_bbox", heat.data_ptr(), float(thr), labels.data_ptr() if want_labels else None,
              counts.data_ptr(), boxes.data_ptr(), int(max_boxes), n, H, W, _stream())
    return labels, counts, boxes
