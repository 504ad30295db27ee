"""Train-step timing for the training-mode processor (SURVEY.md §8 f N3): forward + backward of the SD-1.5 attention
stack (self -> cross per block, UNet weights frozen, the prompt embedding carries the graph, loss = output term + L1 on
the aggregated heat map of three tokens), our processor against a plain-PyTorch restatement of hook.py:83-122 on the
same GPU.  usage: python tools/bench_train.py [batch] [dtype: bf16|fp32]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from agenda_b200 import UNetCrossAttentionHooker
from agenda_b200.sd_attention import AttentionStack, sd15_blocks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dtype = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "fp32") else torch.bfloat16
TOK = [5, 6, 7]


class TorchHooker:
    """hook.py:14-122 in plain torch ops (materialised probabilities), selected tokens only after the head mean."""

    def __init__(self, latent_hw=64):
        self.cross_attn_maps, self.latent_hw = [], latent_hw

    def clear(self):
        self.cross_attn_maps.clear()

    def compute_global_heat_map(self):
        up = [F.interpolate(m.float(), size=(self.latent_hw,) * 2, mode="bicubic").clamp_(min=0) for m in self.cross_attn_maps]
        return torch.stack(up, 0).mean(0)

    def __call__(self, attn, hs, encoder_hidden_states=None, attention_mask=None):
        Bc, N, _ = hs.shape
        H = attn.heads
        q = attn.to_q(hs)
        ctx = hs if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(ctx), attn.to_v(ctx)

        def split(t):
            return t.view(Bc, t.shape[1], H, -1).permute(0, 2, 1, 3).reshape(Bc * H, t.shape[1], -1)
        q, k, v = split(q), split(k), split(v)
        p = torch.baddbmm(torch.empty(Bc * H, N, k.shape[1], dtype=q.dtype, device=q.device), q, k.transpose(-1, -2),
                          beta=0, alpha=attn.scale).softmax(dim=-1)
        if encoder_hidden_states is not None:
            h = w = int(math.sqrt(N))
            self.cross_attn_maps.append(p.view(Bc, H, N, -1).mean(1).permute(0, 2, 1)[:, TOK].reshape(Bc, len(TOK), h, w))
        o = torch.bmm(p, v).view(Bc, H, N, -1).permute(0, 2, 1, 3).reshape(Bc, N, -1)
        return attn.to_out[1](attn.to_out[0](o))


stack = AttentionStack(sd15_blocks(), 768, seed=0).cuda().to(dtype)
for p_ in stack.parameters():
    p_.requires_grad_(False)
hs, ctx0 = stack.make_inputs(B, "cuda", dtype)
tgt = torch.rand(B, len(TOK), 64, 64, device="cuda")


def step(proc):
    proc.clear()
    ctx = ctx0.clone().requires_grad_(True)
    loss = 0.0
    for b, a1, a2 in zip(stack.blocks, stack.attn1, stack.attn2):
        x = hs[(b.hw, b.channels)]
        y = proc(a2, proc(a1, x) + x, ctx)        # self -> cross, residual keeps the scale
        y = proc(a1, y)                           # activations now carry the graph: self-attention backward runs
        loss = loss + y.float().pow(2).mean()
    loss = loss + 50.0 * (proc.compute_global_heat_map() - tgt).abs().mean()
    loss.backward()
    return loss.detach(), ctx.grad


def timed(proc, reps=5):
    for _ in range(2):
        step(proc)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        loss, g = step(proc)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, loss.item(), g, torch.cuda.max_memory_allocated() / 2**30


for m in list(stack.attn1) + list(stack.attn2):
    m.set_processor(None)
ours = UNetCrossAttentionHooker(is_train=True, latent_hw=64, tokens=TOK, precision="bf16" if dtype == torch.bfloat16 else "fp32")
ms_a, la, ga, mem_a = timed(ours)
ms_b, lb, gb, mem_b = timed(TorchHooker(64))
rel = (ga.float() - gb.float()).abs().max().item() / gb.float().abs().max().item()
print(f"train step (48 attention calls fwd+bwd, batch {B}, {dtype}): ours {ms_a:.1f} ms / {mem_a:.1f} GiB peak, "
      f"plain torch {ms_b:.1f} ms / {mem_b:.1f} GiB peak; loss {la:.5f} vs {lb:.5f}; d prompt rel. diff {rel:.2e}")
