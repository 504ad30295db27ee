"""A few launches of the split-precision cross-attention kernel at one shape (for ncu).
usage: python tools/run_x3_once.py [B N H d T launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops

a = [int(x) for x in sys.argv[1:]]
B, N, H, d, T, launches = (a + [16, 4096, 8, 40, 3, 4][len(a):])[:6]
torch.manual_seed(0)
C = H * d
qs = [torch.randn(B, N, C, device="cuda") for _ in range(launches)]
k32 = torch.randn(B, 77, C, device="cuda")
v = torch.randn(B, 77, C, device="cuda").bfloat16()
ctx = ops.pack_context_kv(k32, v, H)
maps = torch.zeros((B // 2, T, N), device="cuda")
for q in qs:
    ops.attn_cross_heat_x3(q, ctx, maps, list(range(5, 5 + T)), B // 2, accumulate=True)
torch.cuda.synchronize()
print("ok", float(maps.sum()))
