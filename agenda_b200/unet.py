"""The rest of the denoising step around the attention processor (SURVEY.md §8 f N4): a self-written SD-1.x UNet
skeleton with random-init weights, a DDIM update with classifier-free guidance, and one CUDA graph per denoising step.

`pipeline(prompt, num_inference_steps=20, generator=...)` at data_generation.py:59 runs, per step, a
UNet2DConditionModel forward at batch 2 x images (CFG) whose 32 attention modules call the installed AttnProcessor
(hook.py:83-122), then the scheduler update.  diffusers is not installable here, so this module restates the SD-1.x
UNet topology (diffusers 0.21.2, from memory: 4 down blocks 320/640/1280/1280 with 2 ResNet layers each, mid block,
4 up blocks with 3 layers each, a Transformer2DModel after every ResNet of the attention-bearing blocks = 16
transformer blocks, H = 8 heads, GEGLU feed-forward, GroupNorm(32), SiLU, sinusoidal time embedding) with plain
PyTorch modules: convolutions / linears / normalisations stay on cuDNN / cuBLAS (library code; not one of the three
subsystems BASELINE.json's north_star rebuilds), every attention call goes through the drop-in processor and therefore
through the hand-written sm_100a kernels.  No VAE and no text encoder: the prompt embedding is synthetic, the output is
the final latent plus the heat maps — the part of the pipeline the heat-map path depends on.

What this gives that the attention stack alone (sd_attention.AttentionStack) does not: hidden states that CHANGE from
step to step and from layer to layer, produced by the real dataflow, so `heat-map-labelled images/s` becomes a
whole-step number (bench.py reports both).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .processor import UNetCrossAttentionHooker
from .sd_attention import SDAttention


def _fast(x: torch.Tensor, groups: int = 32) -> bool:
    return ops.groupnorm_nhwc_supported(x, groups) and not torch.is_grad_enabled()


def group_norm(norm: nn.GroupNorm, x: torch.Tensor, silu: bool = False) -> torch.Tensor:
    """GroupNorm (+ SiLU).  bf16 channels-last activations on the GPU — the pipeline's configuration — go through the
    NHWC kernel (agenda_groupnorm_nhwc): torch's native_group_norm converts a channels-last input to NCHW and the next
    convolution converts it back, which together with the normalisation itself was 28 % of the step
    (profiles/r02_unet_step_torch_profiler.txt).  Anything else (fp32 reference runs, CPU) takes torch's own."""
    if ops.groupnorm_nhwc_supported(x, norm.num_groups) and not torch.is_grad_enabled():
        return ops.groupnorm_nhwc(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu)
    y = norm(x)
    return F.silu(y) if silu else y


def layer_norm(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """LayerNorm; bf16 CUDA activations go through agenda_layernorm (torch's kernel ran at a quarter of the HBM rate on
    these [B*HW, 320..1280] rows: 11 % of the step)."""
    if x.is_cuda and x.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0 and not torch.is_grad_enabled():
        return ops.layernorm(x, norm.weight, norm.bias, norm.eps)
    return norm(x)


class ResnetBlock(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int = 1280):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        if _fast(x, self.norm1.num_groups) and self.conv1.out_channels % (8 * self.norm2.num_groups // math.gcd(8, self.norm2.num_groups)) == 0:
            # inference fast path: the convolutions run without their bias; conv1's bias + the time-embedding projection
            # enter the second GroupNorm as a per-(batch, channel) pre-add, conv2's (+ the shortcut's) bias and the skip
            # connection are one pass — four full-tensor adds fewer per block
            h = F.conv2d(group_norm(self.norm1, x, silu=True), self.conv1.weight, None, padding=1)
            h = h.contiguous(memory_format=torch.channels_last)          # (a no-op after a channels-last cuDNN convolution)
            add = self.time_emb_proj(F.silu(temb)) + self.conv1.bias
            h = ops.groupnorm_nhwc(h, self.norm2.weight, self.norm2.bias, self.norm2.num_groups, self.norm2.eps, True, add)
            h = F.conv2d(h, self.conv2.weight, None, padding=1).contiguous(memory_format=torch.channels_last)
            if self.shortcut is None:
                return ops.add_bias_residual(h, self.conv2.bias, x)
            skip = F.conv2d(x, self.shortcut.weight, None).contiguous(memory_format=torch.channels_last)
            return ops.add_bias_residual(h, self.conv2.bias + self.shortcut.bias, skip)
        h = self.conv1(group_norm(self.norm1, x, silu=True))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(group_norm(self.norm2, h, silu=True))
        return (x if self.shortcut is None else self.shortcut(x)) + h


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        y = self.proj(x)
        if y.is_cuda and y.dtype == torch.bfloat16 and y.shape[-1] % 16 == 0 and not torch.is_grad_enabled():
            return ops.geglu(y)               # one pass over [a | gate] instead of gelu + mul (11 % of the step)
        a, gate = y.chunk(2, dim=-1)
        return a * F.gelu(gate)


class TransformerBlock(nn.Module):
    """Transformer2DModel with one BasicTransformerBlock: GroupNorm -> proj_in -> [LN, attn1, LN, attn2, LN, GEGLU FF]
    -> proj_out + residual.  attn1 / attn2 are SDAttention modules: they call the installed processor."""

    def __init__(self, channels: int, heads: int, context_dim: int, name: str):
        super().__init__()
        self.norm = nn.GroupNorm(32, channels, eps=1e-6)
        self.proj_in = nn.Linear(channels, channels)     # (a 1x1 convolution in SD-1.x: the same map on [B, HW, C])
        self.norm1 = nn.LayerNorm(channels)
        self.attn1 = SDAttention(channels, None, heads, channels // heads)
        self.norm2 = nn.LayerNorm(channels)
        self.attn2 = SDAttention(channels, context_dim, heads, channels // heads)
        self.norm3 = nn.LayerNorm(channels)
        self.ff = nn.Sequential(GEGLU(channels, channels * 4), nn.Linear(channels * 4, channels))
        self.proj_out = nn.Linear(channels, channels)
        self.attn1.block_name = f"{name}.attn1"
        self.attn2.block_name = f"{name}.attn2"

    def forward(self, x, context):
        b, c, h, w = x.shape
        res = x
        t = group_norm(self.norm, x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = self.proj_in(t)
        t = t + self.attn1(layer_norm(self.norm1, t))
        t = t + self.attn2(layer_norm(self.norm2, t), encoder_hidden_states=context)
        t = t + self.ff(layer_norm(self.norm3, t))
        t = self.proj_out(t)
        return t.reshape(b, h, w, c).permute(0, 3, 1, 2) + res


def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """Sinusoidal embedding (flip_sin_to_cos=True, freq_shift=0, as SD-1.x configures it).  t fp32 [B] -> [B, dim]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class SDUNet(nn.Module):
    """SD-1.x UNet2DConditionModel skeleton.  forward(latent [B,4,L,L], t [B] fp32, context [B,77,768]) -> eps [B,4,L,L]."""

    def __init__(self, context_dim: int = 768, channels: Sequence[int] = (320, 640, 1280, 1280), heads: int = 8,
                 seed: int = 0):
        super().__init__()
        c0 = channels[0]
        self.channels = tuple(channels)
        self.context_dim = context_dim
        self.conv_in = nn.Conv2d(4, c0, 3, padding=1)
        self.time_mlp = nn.Sequential(nn.Linear(c0, c0 * 4), nn.SiLU(), nn.Linear(c0 * 4, c0 * 4))
        temb = c0 * 4
        self.down_res, self.down_attn, self.down_sample = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        skip_channels: List[int] = [c0]
        cin = c0
        for level, cout in enumerate(channels):
            with_attn = level < len(channels) - 1
            for i in range(2):
                self.down_res.append(ResnetBlock(cin, cout, temb))
                self.down_attn.append(TransformerBlock(cout, heads, context_dim, f"down{level}.{i}") if with_attn else None)
                cin = cout
                skip_channels.append(cin)
            if level < len(channels) - 1:
                self.down_sample.append(nn.Conv2d(cin, cin, 3, stride=2, padding=1))
                skip_channels.append(cin)
            else:
                self.down_sample.append(None)
        self.mid_res1 = ResnetBlock(cin, cin, temb)
        self.mid_attn = TransformerBlock(cin, heads, context_dim, "mid")
        self.mid_res2 = ResnetBlock(cin, cin, temb)
        self.up_res, self.up_attn, self.up_sample = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        n_levels = len(channels)
        for level in range(n_levels):            # up block `level` mirrors down block n_levels-1-level
            cout = channels[n_levels - 1 - level]
            with_attn = level > 0
            for i in range(3):
                skip = skip_channels.pop()
                self.up_res.append(ResnetBlock(cin + skip, cout, temb))
                self.up_attn.append(TransformerBlock(cout, heads, context_dim, f"up{level}.{i}") if with_attn else None)
                cin = cout
            self.up_sample.append(nn.Conv2d(cin, cin, 3, padding=1) if level < n_levels - 1 else None)
        self.norm_out = nn.GroupNorm(32, cin, eps=1e-5)
        self.conv_out = nn.Conv2d(cin, 4, 3, padding=1)
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():   # deterministic random init, independent of the global RNG state
            for p in self.parameters():
                if p.dim() >= 2:
                    p.copy_(torch.empty_like(p).uniform_(-1, 1, generator=gen) * (p[0].numel() ** -0.5))
            for m in self.modules():
                if isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
                    m.weight.fill_(1.0)
                    m.bias.zero_()
                elif isinstance(m, (nn.Linear, nn.Conv2d)) and m.bias is not None:
                    m.bias.zero_()

    def attention_modules(self):
        return [m for m in self.modules() if isinstance(m, SDAttention)]

    def set_attn_processor(self, processor):
        for m in self.attention_modules():
            m.set_processor(processor)

    def forward(self, latent, t, context):
        temb = self.time_mlp(timestep_embedding(t, self.channels[0]).to(latent.dtype))
        x = self.conv_in(latent)
        skips = [x]
        k = 0
        for level in range(len(self.channels)):
            for _ in range(2):
                x = self.down_res[k](x, temb)
                if self.down_attn[k] is not None:
                    x = self.down_attn[k](x, context)
                skips.append(x)
                k += 1
            if self.down_sample[level] is not None:
                x = self.down_sample[level](x)
                skips.append(x)
        x = self.mid_res2(self.mid_attn(self.mid_res1(x, temb), context), temb)
        k = 0
        for level in range(len(self.channels)):
            for _ in range(3):
                x = self.up_res[k](torch.cat([x, skips.pop()], dim=1), temb)
                if self.up_attn[k] is not None:
                    x = self.up_attn[k](x, context)
                k += 1
            if self.up_sample[level] is not None:
                x = self.up_sample[level](F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return self.conv_out(group_norm(self.norm_out, x, silu=True))


class DDIMSchedule:
    """DDIM (eta = 0) over SD's scaled-linear betas (0.00085 -> 0.012, 1000 training steps), `leading` timestep spacing
    with steps_offset = 1, epsilon prediction — the scheduler arithmetic of one denoising step."""

    def __init__(self, num_inference_steps: int, num_train_steps: int = 1000):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, num_train_steps, dtype=torch.float64) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        ratio = num_train_steps // num_inference_steps
        self.timesteps = [(num_inference_steps - 1 - i) * ratio + 1 for i in range(num_inference_steps)]
        self.ratio = ratio

    def coefficients(self, i: int):
        """x_prev = c_x * x + c_eps * eps for inference step i: (t, c_x, c_eps)."""
        t = self.timesteps[i]
        a_t = float(self.alphas_cumprod[t])
        t_prev = t - self.ratio
        a_prev = float(self.alphas_cumprod[t_prev]) if t_prev >= 0 else float(self.alphas_cumprod[0])
        c_x = math.sqrt(a_prev / a_t)
        c_eps = math.sqrt(1.0 - a_prev) - math.sqrt(a_prev * (1.0 - a_t) / a_t)
        return t, c_x, c_eps


class UNetHeatmapPipeline:
    """`with trace(pipeline): pipeline(prompt, num_inference_steps, generator)` + compute_global_heat_map + the
    post-processing, for a batch of images, on the UNet skeleton: latents from each image's own seed, synthetic prompt
    embeddings, classifier-free guidance, DDIM, the heat-map processor on all 32 attention modules, one CUDA graph for
    the denoising step (timestep-dependent scalars live in device buffers that are refreshed between replays)."""

    def __init__(self, tokens: Sequence[int] = (5, 6, 7), num_steps: int = 50, latent_hw: int = 64, image_size: int = 112,
                 dtype: torch.dtype = torch.bfloat16, device="cuda", guidance_scale: float = 7.5, thr: float = 0.5,
                 max_boxes: int = 64, seed: int = 0, use_cuda_graph: bool = True,
                 channels: Sequence[int] = (320, 640, 1280, 1280), context_dim: int = 768, cross_logits: str = "fp32",
                 cudnn_benchmark: bool = False):
        from .mixed import compensate_cross_projections
        if cudnn_benchmark:
            # process-wide torch switch: cuDNN times its convolution algorithms once per shape (-11 % convolution time here)
            torch.backends.cudnn.benchmark = True
        self.device = torch.device(device)
        self.dtype = dtype
        self.tokens = list(tokens)
        self.num_steps, self.latent_hw, self.image_size = num_steps, latent_hw, image_size
        self.guidance_scale, self.thr, self.max_boxes = guidance_scale, thr, max_boxes
        unet = SDUNet(context_dim, channels, seed=seed)
        if dtype != torch.float32 and cross_logits == "fp32":
            compensate_cross_projections(unet)
        self.unet = unet.to(device=self.device, dtype=dtype).to(memory_format=torch.channels_last)
        self.proc = UNetCrossAttentionHooker(is_train=False, latent_hw=latent_hw, tokens=self.tokens, precision="bf16",
                                             cross_logits=cross_logits)
        self.unet.set_attn_processor(self.proc)
        self.schedule = DDIMSchedule(num_steps)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static = None

    def make_inputs(self, seeds: Sequence[int]):
        """Initial latents [n,4,L,L] and prompt embeddings [2n,77,D] ([uncond..., cond...]) — each image from its own seed."""
        lat, unc, cnd = [], [], []
        D = self.unet.context_dim
        for s in seeds:
            gen = torch.Generator(device=self.device).manual_seed(100003 * int(s) + 29)
            lat.append(torch.randn(1, 4, self.latent_hw, self.latent_hw, generator=gen, device=self.device))
            e = torch.randn(2, 77, D, generator=gen, device=self.device)
            unc.append(e[0:1]); cnd.append(e[1:2])
        return torch.cat(lat, 0).to(self.dtype), torch.cat(unc + cnd, 0).to(self.dtype)

    def _step(self, st):
        """One denoising step on the static buffers: UNet at batch 2n, guidance, DDIM update in place."""
        lat = st["latent"]
        x2 = torch.cat([lat, lat], 0).contiguous(memory_format=torch.channels_last)
        eps = self.unet(x2, st["t"], st["ctx"])
        n = lat.shape[0]
        eps_u, eps_c = eps[:n].float(), eps[n:].float()
        eps_g = eps_u + self.guidance_scale * (eps_c - eps_u)
        lat.copy_((st["coef"][0] * lat.float() + st["coef"][1] * eps_g).to(lat.dtype))

    @torch.no_grad()
    def run(self, latents: torch.Tensor, ctx: torch.Tensor):
        """Returns heat [n,T,L,L] fp32, planes / stack / inv u8, counts, boxes and the final latents."""
        n = latents.shape[0]
        st = self._static
        if st is None or st["latent"].shape != latents.shape or st["ctx"].shape != ctx.shape:
            st = {"latent": torch.empty_like(latents), "ctx": torch.empty_like(ctx),
                  "t": torch.zeros(2 * n, dtype=torch.float32, device=self.device),
                  "coef": torch.zeros(2, dtype=torch.float32, device=self.device)}
            self._static, self._graph = st, None
        st["latent"].copy_(latents)
        st["ctx"].copy_(ctx)
        proc = self.proc
        proc.clear(keep_context_kv=self._graph is not None)
        proc.refresh_context_kv(force=True)
        coefs = [self.schedule.coefficients(i) for i in range(self.num_steps)]

        def set_step(i):
            t, c_x, c_eps = coefs[i]
            st["t"].fill_(float(t))
            st["coef"].copy_(torch.tensor([c_x, c_eps], dtype=torch.float32), non_blocking=True)

        if self.use_cuda_graph and self._graph is None:
            set_step(0)
            saved = st["latent"].clone()
            self._step(st)                        # warm-up outside capture (allocator, cuDNN plans, prompt K/V cache)
            proc.clear(keep_context_kv=True)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step(st)
            self._graph = g
            self._maps_per_step = proc.num_maps
            proc.clear(keep_context_kv=True)
            st["latent"].copy_(saved)
        for i in range(self.num_steps):
            set_step(i)
            if self._graph is not None:
                self._graph.replay()
            else:
                self._step(st)
        if self._graph is not None:
            proc._count = self._maps_per_step * self.num_steps
        heat = proc.compute_global_heat_map()
        planes, stack, inv = ops.heat_postprocess_stack(heat[:, :3].contiguous(), self.image_size)
        _, counts, boxes = ops.ccl_bbox(heat[:, 0].contiguous(), self.thr, self.max_boxes, want_labels=False)
        return {"heat": heat, "planes": planes, "stack": stack, "inv": inv, "counts": counts, "boxes": boxes,
                "latents": st["latent"].clone()}
