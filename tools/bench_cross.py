"""Micro-benchmark of the cross-attention + heat-epilogue kernel.  usage: python tools/bench_cross.py [B N H d] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops

B, N, H, d = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (16, 4096, 8, 40)
T = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.manual_seed(0)
q = torch.randn(B, N, H * d, device="cuda").bfloat16()
k = torch.randn(B, 77, H * d, device="cuda").bfloat16()
v = torch.randn(B, 77, H * d, device="cuda").bfloat16()
toks = None if T >= 77 else list(range(5, 5 + T))
maps = torch.zeros((B // 2, 77 if toks is None else T, N), device="cuda")
# rotate over several Q buffers so the inputs do not sit in L2 between launches
# AGENDA_BENCH_WARM=1: one Q buffer that stays in L2 (the in-pipeline case when Q was just written by the to_q GEMM)
qs = [q] if os.environ.get("AGENDA_BENCH_WARM") == "1" else [q.clone() for _ in range(max(1, int(300e6 // (q.numel() * 2))))]
for _ in range(3):
    ops.attn_cross_heat(q, k, v, H, maps, toks, B // 2, accumulate=True)
torch.cuda.synchronize()
# small layers take less time than a ctypes call: replay the launches from a CUDA graph so the device is timed
reps = 20
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(reps):
        ops.attn_cross_heat(qs[i % len(qs)], k, v, H, maps, toks, B // 2, accumulate=True)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byts = 2 * B * N * H * d * 2 + 2 * B * 77 * H * d * 2 + maps.numel() * 4 * 2
print(f"cross B={B} N={N} H={H} d={d} T={maps.shape[1]}: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s (algorithmic)")
