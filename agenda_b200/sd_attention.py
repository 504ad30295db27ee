"""Host-side stand-ins for the diffusers objects the processor plugs into, plus the synthetic SD-1.x / SD-2.1
attention stacks BASELINE.json's configs are quoted on.

diffusers is not installable here (no network), so `SDAttention` provides the attribute/method surface of
diffusers 0.21.2 `Attention` that an AttnProcessor touches (hook.py:92-120): to_q/to_k/to_v (bias-free), to_out
[Linear, Dropout], heads, scale, norm_cross, prepare_attention_mask, set_processor.  With a real diffusers UNet the
processor is installed unchanged (`unet.set_attn_processor(UNetCrossAttentionHooker(...))`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
from torch import nn


class SDAttention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8,
                 dim_head: int = 64, upcast_attention: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.is_cross = cross_attention_dim is not None
        kv_dim = cross_attention_dim if self.is_cross else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = upcast_attention
        self.upcast_softmax = False
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        """diffusers' contract: an additive mask [B, M], [B, 1 | N, M] or already [B*heads, 1 | N, M] comes back as
        [B*heads, 1 | N, M] (repeated over the heads)."""
        if attention_mask is None:
            return None
        if attention_mask.dim() == 2:
            attention_mask = attention_mask[:, None, :]
        if attention_mask.shape[0] == batch_size:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        if self.processor is None:
            raise RuntimeError("SDAttention has no processor installed")
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)


@dataclass
class BlockSpec:
    name: str
    hw: int      # feature-map side: N = hw*hw query tokens
    channels: int
    heads: int

    @property
    def dim_head(self) -> int:
        return self.channels // self.heads


def sd15_blocks(latent_hw: int = 64) -> List[BlockSpec]:
    """The 16 transformer blocks of the SD-1.x UNet at a `latent_hw`² latent: H=8 everywhere,
    d = C/8 = 40/80/160."""
    L = latent_hw
    spec = []
    for i in range(2): spec.append(BlockSpec(f"down0.{i}", L, 320, 8))
    for i in range(2): spec.append(BlockSpec(f"down1.{i}", L // 2, 640, 8))
    for i in range(2): spec.append(BlockSpec(f"down2.{i}", L // 4, 1280, 8))
    spec.append(BlockSpec("mid", L // 8, 1280, 8))
    for i in range(3): spec.append(BlockSpec(f"up1.{i}", L // 4, 1280, 8))
    for i in range(3): spec.append(BlockSpec(f"up2.{i}", L // 2, 640, 8))
    for i in range(3): spec.append(BlockSpec(f"up3.{i}", L, 320, 8))
    return spec


def sd21_blocks(latent_hw: int = 96) -> List[BlockSpec]:
    """SD-2.1 (768² -> 96² latent): same channels, d = 64 so H = 5/10/20/20."""
    L = latent_hw
    spec = []
    for i in range(2): spec.append(BlockSpec(f"down0.{i}", L, 320, 5))
    for i in range(2): spec.append(BlockSpec(f"down1.{i}", L // 2, 640, 10))
    for i in range(2): spec.append(BlockSpec(f"down2.{i}", L // 4, 1280, 20))
    spec.append(BlockSpec("mid", L // 8, 1280, 20))
    for i in range(3): spec.append(BlockSpec(f"up1.{i}", L // 4, 1280, 20))
    for i in range(3): spec.append(BlockSpec(f"up2.{i}", L // 2, 640, 10))
    for i in range(3): spec.append(BlockSpec(f"up3.{i}", L, 320, 5))
    return spec


class AttentionStack(nn.Module):
    """attn1 (self) + attn2 (cross) of every transformer block, i.e. the 32 processor calls of one UNet forward
    (SURVEY.md §3.1).  The non-attention UNet layers are out of scope (SURVEY.md §8 f N4): each block is fed a
    fixed synthetic hidden state of its own shape."""

    def __init__(self, blocks: List[BlockSpec], context_dim: int = 768, seed: int = 0):
        super().__init__()
        self.blocks = blocks
        self.context_dim = context_dim
        gen = torch.Generator().manual_seed(seed)
        self.attn1 = nn.ModuleList()
        self.attn2 = nn.ModuleList()
        for b in blocks:
            self.attn1.append(SDAttention(b.channels, None, b.heads, b.dim_head))
            self.attn2.append(SDAttention(b.channels, context_dim, b.heads, b.dim_head))
            # diffusers-style qualified names ("mid_block....attn2") for code that selects layers by name
            self.attn1[-1].block_name = f"{b.name}.attn1"
            self.attn2[-1].block_name = f"{b.name}.attn2"
        with torch.no_grad():
            for p in self.parameters():  # deterministic random init, independent of global RNG state
                p.copy_(torch.empty_like(p).uniform_(-1, 1, generator=gen) * (p.shape[-1] ** -0.5))

    def set_attn_processor(self, processor):
        for m in list(self.attn1) + list(self.attn2):
            m.set_processor(processor)

    def make_inputs(self, batch: int, device, dtype, seed: int = 0, context_len: int = 77):
        gen = torch.Generator().manual_seed(1000 + seed)
        hs = {}
        for b in self.blocks:
            key = (b.hw, b.channels)
            if key not in hs:
                hs[key] = torch.randn(batch, b.hw * b.hw, b.channels, generator=gen).to(device=device, dtype=dtype)
        ctx = torch.randn(batch, context_len, self.context_dim, generator=gen).to(device=device, dtype=dtype)
        return hs, ctx

    def make_inputs_for_seeds(self, seeds, device, dtype, context_len: int = 77, on_device: bool = False):
        """Synthetic inputs of a batch of images, each a function of ITS OWN seed only (data_generation.py:56-59: an image
        depends on (prompt, seed), never on which other seeds share its batch or rank).  UNet batch layout as under
        classifier-free guidance: [uncond_0 .. uncond_{n-1}, cond_0 .. cond_{n-1}] (hook.py:48-49 keeps the second half).
        on_device=True draws the values with a per-image generator ON `device` (no host synthesis, no H2D copy; the
        values differ from the CPU generator's, but are again a function of the seed alone)."""
        shapes = []
        for b in self.blocks:
            if (b.hw, b.channels) not in shapes:
                shapes.append((b.hw, b.channels))
        gdev = device if on_device else "cpu"
        per_image = []
        for s in seeds:
            gen = torch.Generator(device=gdev).manual_seed(100003 * int(s) + 17)
            item = {k: torch.randn(2, k[0] * k[0], k[1], generator=gen, device=gdev) for k in shapes}
            item["ctx"] = torch.randn(2, context_len, self.context_dim, generator=gen, device=gdev)
            per_image.append(item)

        def batch_of(key):
            t = torch.cat([im[key][0:1] for im in per_image] + [im[key][1:2] for im in per_image], 0)
            return t.to(device=device, dtype=dtype)
        return {k: batch_of(k) for k in shapes}, batch_of("ctx")

    def forward(self, hidden_by_shape, context):
        """One UNet forward's worth of attention calls; returns the last output (keeps the work observable)."""
        out = None
        for b, a1, a2 in zip(self.blocks, self.attn1, self.attn2):
            hs = hidden_by_shape[(b.hw, b.channels)]
            out = a1(hs)
            out = a2(hs, encoder_hidden_states=context)
        return out
