"""Top CUDA kernels of one eager train step of the training-mode processor (torch.profiler): the step of
`bench.py --workload train`.  usage: python tools/profile_train.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from agenda_b200 import UNetCrossAttentionHooker
from agenda_b200.sd_attention import AttentionStack, sd15_blocks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TOK = [5, 6, 7]
dev = torch.device("cuda")
stack = AttentionStack(sd15_blocks(), 768, seed=0).to(dev).to(torch.bfloat16)
for p_ in stack.parameters():
    p_.requires_grad_(False)
hs, ctx0 = stack.make_inputs(B, dev, torch.bfloat16)
tgt = torch.rand(B, len(TOK), 64, 64, device=dev)
proc = UNetCrossAttentionHooker(is_train=True, latent_hw=64, tokens=TOK, precision="bf16")


def step():
    proc.clear()
    ctx = ctx0.clone().requires_grad_(True)
    loss = 0.0
    for b, a1, a2 in zip(stack.blocks, stack.attn1, stack.attn2):
        x = hs[(b.hw, b.channels)]
        y = proc(a2, proc(a1, x) + x, ctx)
        y = proc(a1, y)
        loss = loss + y.float().pow(2).mean()
    loss = loss + 50.0 * (proc.compute_global_heat_map() - tgt).abs().mean()
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=70))
