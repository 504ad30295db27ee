"""Micro-benchmark of agenda_attn_cross_bwd (cross-attention + heat backward).  usage: python tools/bench_cross_bwd.py [B N H d] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops

B, N, H, d = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (2, 4096, 8, 40)
T = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.manual_seed(0)
q = torch.randn(B, N, H * d, device="cuda").bfloat16()
k = torch.randn(B, 77, H * d, device="cuda").bfloat16()
v = torch.randn(B, 77, H * d, device="cuda").bfloat16()
go = torch.randn_like(q)
gm = torch.randn(B, T, N, device="cuda")
toks = list(range(5, 5 + T))
for _ in range(3):
    ops.attn_cross_bwd(q, k, v, go, gm, H, toks, 0)
torch.cuda.synchronize()
reps = 20
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(reps):
        ops.attn_cross_bwd(q, k, v, go, gm, H, toks, 0)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3          # includes the two memsets and two casts of dk / dv (small)
byts = 3 * B * N * H * d * 2 + 2 * B * 77 * H * d * 2 + 2 * B * 77 * H * d * 4 + gm.numel() * 4
flops = 10.0 * B * H * N * 77 * d
print(f"cross bwd B={B} N={N} H={H} d={d} T={T}: {us:.1f} us  {byts / us / 1e3:.0f} GB/s (algorithmic)  {flops / us / 1e6:.1f} TFLOP/s fp32")
