"""Kernel times of one cross-attention backward call (torch.profiler).  usage: python tools/profile_cross_bwd.py [B N H d T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from agenda_b200 import ops

B, N, H, d, T = (int(a) for a in sys.argv[1:6]) if len(sys.argv) > 5 else (2, 4096, 8, 40, 3)
dev = "cuda"
q = torch.randn(B, N, H * d, device=dev).bfloat16()
k = torch.randn(B, 77, H * d, device=dev).bfloat16()
v = torch.randn(B, 77, H * d, device=dev).bfloat16()
go = torch.randn_like(q)
gm = torch.randn(B, T, N, device=dev)
tok = list(range(5, 5 + T))
for _ in range(3):
    ops.attn_cross_bwd(q, k, v, go, gm, H, tok, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.attn_cross_bwd(q, k, v, go, gm, H, tok, 0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=12, max_name_column_width=90))
