"""Drop-in for the reference's `UNetCrossAttentionHooker` (data_generation/hook.py:14-122).

Same plug-in surface — a diffusers AttnProcessor callable
    proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None) -> Tensor[B,N,C]
installed with `unet.set_attn_processor(proc)` (finetune_sd_token.py:755-757) — and the same extra surface
(`cross_attn_maps`, `is_train`, `latent_hw`, `clear()`, `compute_global_heat_map()`), but the work happens in
hand-written sm_100a kernels behind the C ABI:

  * hook.py:104-115 (head_to_batch_dim, baddbmm, softmax, bmm, batch_to_head_dim) -> one fused attention kernel;
    the [B*H,N,M] probability tensor is never materialised.
  * hook.py:28-56 (`_unravel_attn`: permute, 77-iteration Python loop, stack, mean over heads) -> the cross-attention
    kernel's epilogue: selected-token probabilities are summed over heads on chip and leave as [B',T,N] fp32.
  * hook.py:59-81 (`compute_global_heat_map`: bicubic -> clamp -> stack -> mean) -> streaming: every call adds
    clamp(bicubic(map)) into one persistent [B',T,L,L] fp32 buffer (directly from the attention epilogue when the
    layer is already at latent resolution), `compute_global_heat_map()` divides by the number of maps.

Differences from the reference, by design:
  * `tokens=[...]` restricts the heat maps to the listed key-token rows (the reference always keeps all 77 and its
    callers then read a handful, data_generation.py:74-77).  `tokens=None` keeps all of them, as the reference does.
  * `cross_attn_maps` is only populated when `record_maps=True` (the reference's list costs 415 MB/image at 50
    steps); aggregation never needs it.
  * training (`is_train=True` under autograd, finetune_sd_token.py:1043-1069; SURVEY.md §8 f N3): when any input of
    a call requires grad, the call goes through `agenda_b200/autograd.py` — forward in the same CUDA kernels,
    backward by recomputation with library kernels (first correct version) — every cross-attention map [B',T,h,w] is
    appended to `cross_attn_maps` with its graph like the reference's list, and `compute_global_heat_map()` aggregates
    them with differentiable torch ops (hook.py:59-81).
  * `aggregate="daam"` switches the aggregation to the `daam` package's (what data_generation.py:57-77 actually
    calls; un-vendored, parity unpinned — SURVEY.md §8 a5): per hooked layer one [B',H,T,h,w] buffer sums the per-head
    probabilities over the denoising steps at native resolution (the attention epilogue adds into it), layers whose
    map is latent_hw/8 wide are skipped, and `compute_global_heat_map()` upsamples + clamps every (layer, head) plane
    on its own before averaging over all of them.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import torch

from . import autograd as _ag
from . import ops


_ADDMM_IN_PLACE = None   # does this torch build take addmm(out_dtype=fp32, out=input) on 16-bit operands?


def _addmm_f32_(q: torch.Tensor, x2: torch.Tensor, w_t: torch.Tensor) -> torch.Tensor:
    """q (fp32) += x2 @ w_t (16-bit operands, fp32 accumulation) — as ONE GEMM that accumulates into q when the
    library allows it, else a second fp32-output GEMM and an add."""
    global _ADDMM_IN_PLACE
    if _ADDMM_IN_PLACE is not False:
        try:
            torch.addmm(q, x2, w_t, out_dtype=torch.float32, out=q)
            _ADDMM_IN_PLACE = True
            return q
        except (RuntimeError, TypeError):
            if _ADDMM_IN_PLACE:   # worked before: a real error
                raise
            _ADDMM_IN_PLACE = False
    q += torch.mm(x2, w_t, out_dtype=torch.float32)
    return q


class UNetCrossAttentionHooker:
    def __init__(self, is_train: bool = True, latent_hw: int = 64, tokens: Optional[Sequence[int]] = None,
                 precision: str = "bf16", record_maps: bool = False, aggregate: str = "hook",
                 cross_logits: str = "fp32"):
        if aggregate not in ("hook", "daam"):
            raise ValueError("aggregate must be 'hook' or 'daam'")
        if cross_logits not in ("fp32", "bf16"):
            raise ValueError("cross_logits must be 'fp32' (split-precision tensor-core kernel) or 'bf16'")
        self.aggregate = aggregate
        # precision="bf16" runs the attention cores on the tensor cores.  For cross-attention, whose probabilities ARE
        # the product (the heat maps, held to 1e-4 against the reference's fp32 run), cross_logits="fp32" keeps the
        # logits at fp32 accuracy: to_q with fp32 accumulation AND fp32 output, the key projection in fp32, and the
        # split-precision kernel (agenda_attn_cross_fwd_heat_x3: Q_hi K_hi + Q_lo K_hi + Q_hi K_lo on bf16 tensor
        # cores).  cross_logits="bf16" rounds Q and K to bf16 first (faster, heat maps to ~2e-3 only).
        self.cross_logits = "bf16" if os.environ.get("AGENDA_CROSS_X3", "1") == "0" else cross_logits
        self._layer_sums = {}   # aggregate="daam": id(attn) -> [B',H,T,h,w] fp32 sum over denoising steps
        self._layer_mods = {}   # id(attn) -> attn (held so that a freed module's id cannot be reused for another)
        # self-attention, bf16: to_q/to_k/to_v (hook.py:93,101-102) act on the same hidden_states, so they run as ONE
        # GEMM with the concatenated weights (hidden_states read once instead of three times) and the attention kernel
        # reads q/k/v as column slices of its output.  id(attn) -> (weight versions, fused [3C,C] weight)
        self.fuse_qkv = os.environ.get("AGENDA_FUSE_QKV", "1") != "0"
        self.prescale_q = os.environ.get("AGENDA_PRESCALE_Q", "1") != "0"
        self._qkv_weights = {}
        # cross-attention: to_k / to_v of the prompt embedding (hook.py:101-102) do not depend on the latent, yet the
        # reference recomputes them in every block at every denoising step.  K and V are kept per module for as long as
        # the SAME encoder_hidden_states tensor object (held here, so its storage cannot be recycled) is passed again
        # unmodified (`_version`) and the two weights are unchanged.  id(attn) -> _ContextKV
        self.cache_context_kv = os.environ.get("AGENDA_CACHE_CONTEXT_KV", "1") != "0"
        self._ctx_kv = {}
        self.cross_attn_maps: List[torch.Tensor] = []
        self.is_train = is_train
        self.latent_hw = latent_hw
        self.tokens = None if tokens is None else [int(t) for t in tokens]
        self.precision = precision
        self.record_maps = record_maps
        self._acc: Optional[torch.Tensor] = None  # [B', T, L, L] fp32 running sum of clamp(bicubic(map))
        self._count = 0

    # ---- hook.py:25-26 -------------------------------------------------------------------------------------
    def clear(self, keep_context_kv: bool = False):
        """hook.py:25-26.  Also forgets the cached prompt K/V unless `keep_context_kv` (a caller that replays a
        captured CUDA graph reading those buffers keeps them and calls refresh_context_kv())."""
        self.cross_attn_maps.clear()
        if not keep_context_kv:
            self._ctx_kv.clear()
        if self._acc is not None:
            self._acc.zero_()
        for buf in self._layer_sums.values():
            buf.zero_()
        self._count = 0

    @property
    def num_maps(self) -> int:
        return self._count

    # ---- hook.py:59-81 -------------------------------------------------------------------------------------
    def compute_global_heat_map(self) -> torch.Tensor:
        """[B', T, latent_hw, latent_hw] fp32 (T = 77 rows, or len(tokens))."""
        if self.aggregate == "daam":
            if self._count == 0 or not self._layer_sums:
                raise RuntimeError('No heat maps found.')
            first = next(iter(self._layer_sums.values()))
            L = self.latent_hw
            acc = torch.zeros((first.shape[0], first.shape[2], L, L), dtype=torch.float32, device=first.device)
            pairs = 0
            for buf in self._layer_sums.values():   # insertion order == network order: deterministic sum
                ops.heat_upsample_accum_heads(buf, acc)
                pairs += buf.shape[1]
            return ops.heat_finalize(acc, pairs)    # mean over every (layer, head) map
        if any(m.requires_grad for m in self.cross_attn_maps):   # training: differentiable aggregation
            if self._count != len(self.cross_attn_maps):
                raise RuntimeError("heat maps with and without an autograd graph were mixed; call clear() between "
                                   "a no-grad pass and a training pass")
            return _ag.global_heat_map_autograd(self.cross_attn_maps, self.latent_hw)
        if self._count == 0 or self._acc is None:
            raise RuntimeError('No heat maps found.')
        return ops.heat_finalize(self._acc, self._count)

    def _fused_qkv_weight(self, attn):
        """[3C, C] concatenation of to_q/to_k/to_v weights (bias-free Linear layers of equal shape), cached per module
        and rebuilt when any of the three weights is modified in place or replaced."""
        mods = (attn.to_q, attn.to_k, attn.to_v)
        ws = [getattr(m, "weight", None) for m in mods]
        if any(w is None or w.dtype != torch.bfloat16 or not w.is_cuda for w in ws):
            return None, False
        if any(getattr(m, "bias", None) is not None for m in mods) or not (ws[0].shape == ws[1].shape == ws[2].shape):
            return None, False
        if type(attn.to_q) is not torch.nn.Linear or ws[0].shape[0] % 8:
            return None, False
        # d = 40 layers: scale * log2(e) is folded into the q rows (one rounding of the scaled fp32 weight to bf16), so the
        # attention kernel's scores are the base-2 exponents themselves (ABI: scale == 0)
        prescale = self.prescale_q and ws[0].shape[0] // attn.heads == 40
        key = tuple((w.data_ptr(), 0 if w.is_inference() else w._version) for w in ws) + (prescale, float(attn.scale))
        hit = self._qkv_weights.get(id(attn))
        if hit is None or hit[0] != key:
            wq = ws[0].detach()
            if prescale:
                wq = (wq.float() * (float(attn.scale) * 1.4426950408889634)).to(torch.bfloat16)
            # (the module is kept in the entry: a freed module's id() could otherwise be reused by another one)
            hit = (key, torch.cat([wq, ws[1].detach(), ws[2].detach()], 0).contiguous(), prescale, attn)
            self._qkv_weights[id(attn)] = hit
        return hit[1], hit[2]

    @staticmethod
    def _kv_state(attn, ehs):
        wk, wv = attn.to_k.weight, attn.to_v.weight
        # (inference tensors carry no version counter: they cannot be modified in place outside inference mode either)
        ver = 0 if ehs.is_inference() else ehs._version
        return (ver, tuple(ehs.shape), wk.data_ptr(), 0 if wk.is_inference() else wk._version, wv.data_ptr(),
                0 if wv.is_inference() else wv._version)

    @staticmethod
    def _full_weight(lin):
        """fp32 weight of a projection; modules prepared by mixed.compensate_cross_projections carry the bf16 residual
        `weight_lo` of their fp32 checkpoint value beside the bf16 `weight` (W = weight + weight_lo to 2^-17)."""
        w = lin.weight.detach().float()
        lo = getattr(lin, "weight_lo", None)
        return w if lo is None else w + lo.float()

    def _project_context(self, attn, ehs, split: bool, into=None):
        """to_k / to_v of the prompt embedding (hook.py:101-102).  split=False: (K, V) in the modules' own precision.
        split=True (cross_logits="fp32"): K in fp32 — from the fp32 weight, or weight + weight_lo — and V, packed into
        the tensor-core operand blob of the split-precision kernel (ops.pack_context_kv): (ContextKV,).  `into` refills
        an existing blob in place."""
        if not split:
            return attn.to_k(ehs), attn.to_v(ehs)
        k32 = torch.nn.functional.linear(ehs.float(), self._full_weight(attn.to_k))
        bias = getattr(attn.to_k, "bias", None)
        if bias is not None:
            k32 = k32 + bias.float()
        return (ops.pack_context_kv(k32, attn.to_v(ehs), attn.heads, out=into),)

    def _context_kv(self, attn, ehs, split: bool = False):
        if torch.is_grad_enabled() and (ehs.requires_grad or attn.to_k.weight.requires_grad
                                        or attn.to_v.weight.requires_grad):
            return self._project_context(attn, ehs, split)  # never cache tensors that carry an autograd graph
        ent = self._ctx_kv.get(id(attn))
        state = self._kv_state(attn, ehs) + (split,)
        if ent is not None and ent[0] is ehs and ent[1] == state:
            return ent[2]
        tensors = self._project_context(attn, ehs, split)
        self._ctx_kv[id(attn)] = (ehs, state, tensors, attn)
        return tensors

    def refresh_context_kv(self, force: bool = False) -> None:
        """Recompute cached K / V IN PLACE from the tensor they were built from (only those whose source or weights
        changed, or all of them with `force`).  For callers that replay a captured CUDA graph of the processor calls
        after overwriting the prompt embedding in place (HeatmapPipeline): the graph reads the cached buffers, so they
        must be brought up to date outside the graph before the replays."""
        for mod_id, (ehs, state, tensors, attn) in list(self._ctx_kv.items()):
            new_state = self._kv_state(attn, ehs) + (state[-1],)
            if new_state == state and not force:
                continue
            if new_state[1] != state[1]:   # shape changed: rebuilt on the next call
                del self._ctx_kv[mod_id]
                continue
            if state[-1]:
                self._project_context(attn, ehs, True, into=tensors[0])
            else:
                for dst, src in zip(tensors, self._project_context(attn, ehs, False)):
                    dst.copy_(src)
            self._ctx_kv[mod_id] = (ehs, new_state, tensors, attn)

    @staticmethod
    def _query_fp32(attn, hidden_states, chunk_major: bool = False):
        """to_q (hook.py:93) with an fp32 result.  fp32 activations: the module's own fp32 GEMM (what the reference
        runs).  bf16 activations: agenda_linear_split_f32 — one tcgen05 GEMM with fp32 accumulation and fp32 OUTPUT
        (products of 16-bit values are exact, so this is the fp32 projection of those activations and weights) that
        also takes the `weight_lo` residual of the checkpoint's fp32 weight as a second B operand of the same
        accumulator.  Other 16-bit cases (fp16, odd shapes): library GEMMs with out_dtype=float32 (+ a correction GEMM)."""
        lin = attn.to_q
        w = getattr(lin, "weight", None)
        if (hidden_states.dtype == torch.float32 or w is None or w.dtype != hidden_states.dtype
                or type(lin) is not torch.nn.Linear):
            return lin(hidden_states).float()
        B, N, C = hidden_states.shape
        lo = getattr(lin, "weight_lo", None)
        if lin.bias is None and ops.linear_split_f32_supported(hidden_states, w) and (lo is None or lo.dtype == w.dtype):
            # one tcgen05 GEMM: both weight halves into the same fp32 accumulator, activations read once, fp32 written once
            if chunk_major and (w.shape[0] // attn.heads) % 40 == 0 and w.shape[0] % attn.heads == 0:
                # ... in the chunk-major layout the cross-attention kernel fetches with one bulk copy per 128-query chunk
                return ops.linear_split_f32_heads(hidden_states, w.detach(), None if lo is None else lo, attn.heads)
            if os.environ.get("AGENDA_PACKED_TOQ", "0") == "1":
                # opt-in: the weights packed once per (weight, version) into the GEMM's stage images, so that a K block's
                # hi | lo tiles arrive with one bulk copy.  Bit-identical; measured 1-5 % faster only:
                # the kernel is bound by the bytes it pulls into the SM, not by the TMA engine's per-row cost
                try:
                    key = (w.data_ptr(), w._version, None if lo is None else (lo.data_ptr(), lo._version))
                except RuntimeError:      # inference tensors do not track versions
                    key = None
                if key is not None:
                    cached = getattr(lin, "_agenda_packed_w", None)
                    if cached is None or cached[0] != key:
                        cached = (key, ops.linear_split_pack(w.detach(), None if lo is None else lo))
                        lin._agenda_packed_w = cached
                    return ops.linear_split_f32_packed(hidden_states, cached[1])
            return ops.linear_split_f32(hidden_states, w.detach(), None if lo is None else lo)
        x2 = hidden_states.reshape(B * N, C)
        q = torch.mm(x2, w.t(), out_dtype=torch.float32)
        if lo is not None:
            _addmm_f32_(q, x2, lo.to(x2.dtype).t())
        if lin.bias is not None:
            q += lin.bias.float()
        return q.view(B, N, -1)

    def _use_x3(self, attn, query_dim: int, M: int) -> bool:
        if self.precision != "bf16" or self.cross_logits != "fp32":
            return False
        n_tok = M if self.tokens is None else len(self.tokens)
        if n_tok > 8 and self.aggregate == "daam":   # per-head planes: the kernel keeps at most 8 tokens apart
            return False
        return M <= 80 and (query_dim // attn.heads) in (40, 64, 80, 160) and query_dim % attn.heads == 0

    def _accumulate(self, b_kept: int, n_tok: int, device) -> torch.Tensor:
        L = self.latent_hw
        if self._acc is None or self._acc.shape != (b_kept, n_tok, L, L) or self._acc.device != device:
            if self._count:
                raise RuntimeError("heat-map batch/token shape changed between calls; call clear() first")
            self._acc = torch.zeros((b_kept, n_tok, L, L), dtype=torch.float32, device=device)
        return self._acc

    @staticmethod
    def _wants_grad(attn, hidden_states, encoder_hidden_states) -> bool:
        ts = [hidden_states, encoder_hidden_states]
        for m in (attn.to_q, attn.to_k, attn.to_v):
            ts.extend(m.parameters())
        return any(t is not None and t.requires_grad for t in ts)

    def _call_with_autograd(self, attn, hidden_states, encoder_hidden_states):
        """hook.py:83-122 with a graph (finetune_sd_token.py:1043-1069): projections through their own modules, the
        attention cores through autograd.Functions whose forward is the CUDA kernel."""
        if self.aggregate != "hook":
            raise NotImplementedError("agenda_b200: autograd is implemented for aggregate='hook' (the processor the "
                                      "reference trains with), not for the DAAM aggregation")
        batch_size, sequence_length, _ = hidden_states.shape
        query = attn.to_q(hidden_states)
        is_cross_attn = encoder_hidden_states is not None
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        elif attn.norm_cross is not None:
            encoder_hidden_states = attn.norm_cross(encoder_hidden_states)
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        heads, scale = attn.heads, float(attn.scale)
        if is_cross_attn:
            b_first = 0 if self.is_train else batch_size // 2  # hook.py:48-49
            h = w = int(math.sqrt(sequence_length))
            hidden_states, maps = _ag.CrossAttentionHeatFn.apply(query, key, value, heads, scale, self.tokens, b_first)
            self.cross_attn_maps.append(maps.view(maps.shape[0], maps.shape[1], h, w))  # hook.py:110-112
            self._count += 1
        else:
            hidden_states = _ag.SelfAttentionFn.apply(query, key, value, heads, scale, self.precision)
        hidden_states = attn.to_out[0](hidden_states)
        return attn.to_out[1](hidden_states)

    # ---- hook.py:83-122 ------------------------------------------------------------------------------------
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        batch_size, sequence_length, _ = hidden_states.shape
        # hook.py:92.  The SD UNets pass no mask; when one arrives (additive, [B*heads, 1 | N, M]) the exact fp32 kernel
        # adds it to the scaled logits like get_attention_scores' baddbmm (hook.py:108).  `upcast_attention` /
        # `upcast_softmax` need no branch: every kernel here accumulates the logits and takes the softmax in fp32.
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
        if torch.is_grad_enabled() and self._wants_grad(attn, hidden_states, encoder_hidden_states):
            if attention_mask is not None:
                raise NotImplementedError("agenda_b200: attention masks are not supported in training mode")
            return self._call_with_autograd(attn, hidden_states, encoder_hidden_states)
        if (encoder_hidden_states is None and attention_mask is None and self.fuse_qkv and self.precision == "bf16"
                and hidden_states.dtype == torch.bfloat16 and hidden_states.is_cuda):
            w, prescaled = self._fused_qkv_weight(attn)
            if w is not None:
                qkv = torch.nn.functional.linear(hidden_states, w)
                hidden_states = ops.attn_self_fused_qkv(qkv, attn.heads, scale=float(attn.scale), prescaled=prescaled)
                hidden_states = attn.to_out[0](hidden_states)
                return attn.to_out[1](hidden_states)
        is_cross_attn = encoder_hidden_states is not None
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        elif attn.norm_cross is not None:
            encoder_hidden_states = attn.norm_cross(encoder_hidden_states)
        heads = attn.heads
        scale = float(attn.scale)
        in_dtype = hidden_states.dtype

        x3 = (is_cross_attn and attention_mask is None and hidden_states.is_cuda
              and self._use_x3(attn, attn.to_q.weight.shape[0], encoder_hidden_states.shape[1]))
        if x3:
            # AGENDA_Q_CHUNKS=1: to_q writes the chunk-major layout and the kernel fetches Q by bulk copy.  Measured (DESIGN.md
            # K2x3): bit-identical, the attention kernel gains 1.3 us of 48 at 64^2 but the GEMM's direct-store epilogue loses
            # 9.5 us against its staged tensor-map stores, so the row-major form stays the default
            query = self._query_fp32(attn, hidden_states, chunk_major=os.environ.get("AGENDA_Q_CHUNKS", "0") == "1")
            if self.cache_context_kv and attn.norm_cross is None:
                kv = self._context_kv(attn, encoder_hidden_states, split=True)
            else:
                kv = self._project_context(attn, encoder_hidden_states, True)
            out_dtype = in_dtype if in_dtype in (torch.bfloat16, torch.float32) else torch.bfloat16

            def cross(maps, accumulate, per_head=False):
                o = ops.attn_cross_heat_x3(query, kv[0], maps, self.tokens, b_first, accumulate=accumulate, scale=scale,
                                           per_head=per_head, out_dtype=out_dtype)
                return o if o.dtype == in_dtype else o.to(in_dtype)
        else:
            query = attn.to_q(hidden_states)
            if is_cross_attn and self.cache_context_kv and attn.norm_cross is None:
                key, value = self._context_kv(attn, encoder_hidden_states)
            else:
                key = attn.to_k(encoder_hidden_states)
                value = attn.to_v(encoder_hidden_states)

            def cross(maps, accumulate, per_head=False):
                if attention_mask is not None:
                    return ops.attn_masked(query, key, value, heads, attention_mask, maps, self.tokens, b_first,
                                           accumulate=accumulate, scale=scale, per_head=per_head)
                return ops.attn_cross_heat(query, key, value, heads, maps, self.tokens, b_first, accumulate=accumulate,
                                           scale=scale, per_head=per_head)

        if is_cross_attn:
            M = encoder_hidden_states.shape[1]
            n_tok = M if self.tokens is None else len(self.tokens)
            b_first = 0 if self.is_train else batch_size // 2  # hook.py:48-49: drop the unconditional half
            h = w = int(math.sqrt(sequence_length))
            dev = hidden_states.device
            if self.aggregate == "daam":
                if self.latent_hw // h == 8:  # daam drops the factor-8 maps
                    hidden_states = cross(None, False)
                else:
                    buf = self._layer_sums.get(id(attn))
                    shape = (batch_size - b_first, heads, n_tok, h, w)
                    if buf is None or tuple(buf.shape) != shape or buf.device != dev:
                        buf = torch.zeros(shape, dtype=torch.float32, device=dev)
                        self._layer_sums[id(attn)] = buf
                        self._layer_mods[id(attn)] = attn   # keeps the module alive: its id() stays unique
                    hidden_states = cross(buf, True, per_head=True)
                    self._count += 1
                hidden_states = attn.to_out[0](hidden_states)
                return attn.to_out[1](hidden_states)
            acc = self._accumulate(batch_size - b_first, n_tok, dev)
            if h == self.latent_hw and not self.record_maps:
                # bicubic at scale 1 is the identity and probabilities are >= 0: accumulate from the epilogue
                hidden_states = cross(acc, True)
            else:
                maps = torch.empty((batch_size - b_first, n_tok, h, w), dtype=torch.float32, device=dev)
                hidden_states = cross(maps, False)
                ops.heat_upsample_accum(maps, acc)
                if self.record_maps:
                    self.cross_attn_maps.append(maps)
            self._count += 1
        elif attention_mask is not None:
            hidden_states = ops.attn_masked(query, key, value, heads, attention_mask, scale=scale)
        else:
            hidden_states = ops.attn_self(query, key, value, heads, scale=scale, precision=self.precision)

        # linear proj
        hidden_states = attn.to_out[0](hidden_states)
        # dropout
        hidden_states = attn.to_out[1](hidden_states)
        return hidden_states


# descriptive alias used in the docs
B200CrossAttnProcessor = UNetCrossAttentionHooker
