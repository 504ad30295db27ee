"""Micro-benchmark of the self-attention kernel variants (CUDA events, inputs larger than nothing: L2-resident K/V is
the real case).  usage: python tools/bench_attn.py [B N H d] [variants...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import _lib

B, N, H, d = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (16, 4096, 8, 40)
variants = [int(x) for x in sys.argv[5:]] or [0, 1, 2]
torch.manual_seed(0)
q, k, v = (torch.randn(B, N, H * d, device="cuda").bfloat16() for _ in range(3))
out = torch.empty_like(q)
st = torch.cuda.current_stream().cuda_stream
flops = 4.0 * B * H * N * N * d
for var in variants:
    for _ in range(3):
        _lib.call("agenda_attn_self_fwd_variant", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, N, d,
                  float(d ** -0.5), var, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        _lib.call("agenda_attn_self_fwd_variant", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, N, d,
                  float(d ** -0.5), var, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"variant {var}: B={B} N={N} H={H} d={d}  {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s (useful)")
