"""Micro-benchmark: split-precision cross-attention (agenda_attn_cross_fwd_heat_x3, fp32 Q) against the plain bf16
tensor-core kernel on the SD-1.5 / SD-2.1 layer shapes.  usage: python tools/bench_cross_x3.py [images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops

images = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SHAPES = [("sd15 64^2", 2 * images, 4096, 8, 40), ("sd15 32^2", 2 * images, 1024, 8, 80), ("sd15 16^2", 2 * images, 256, 8, 160),
          ("sd15 8^2", 2 * images, 64, 8, 160), ("sd21 96^2", 4 * images, 9216, 5, 64), ("sd21 48^2", 4 * images, 2304, 10, 64)]
T = 3


def timed(fn, reps=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, B, N, H, d in SHAPES:
    torch.manual_seed(0)
    C = H * d
    q32 = torch.randn(B, N, C, device="cuda")
    k32 = torch.randn(B, 77, C, device="cuda")
    v = torch.randn(B, 77, C, device="cuda").bfloat16()
    ctx = ops.pack_context_kv(k32, v, H)
    toks = list(range(5, 5 + T))
    maps = torch.zeros((B // 2, T, N), device="cuda")
    nbuf = 1 if os.environ.get('AGENDA_BENCH_WARM') == '1' else max(1, int(400e6 // (q32.numel() * 4)))   # rotate over Q buffers larger than L2 in total (WARM=1: one buffer)
    q32s = [q32.clone() for _ in range(nbuf)]
    qbs = [x.bfloat16() for x in q32s]
    kb = k32.bfloat16()
    ms_x3 = timed(lambda i: ops.attn_cross_heat_x3(q32s[i % nbuf], ctx, maps, toks, B // 2, accumulate=True))
    ms_bf = timed(lambda i: ops.attn_cross_heat(qbs[i % nbuf], kb, v, H, maps, toks, B // 2, accumulate=True))
    by_x3 = B * N * C * (4 + 2) + 3 * B * 77 * C * 2 + maps.numel() * 4
    by_bf = B * N * C * (2 + 2) + 2 * B * 77 * C * 2 + maps.numel() * 4
    print(f"{name:10s} B={B:3d} N={N:5d} H={H:2d} d={d:3d}: x3 {ms_x3 * 1e3:7.1f} us {by_x3 / ms_x3 / 1e6:6.0f} GB/s | "
          f"bf16 {ms_bf * 1e3:7.1f} us {by_bf / ms_bf / 1e6:6.0f} GB/s", flush=True)
    del q32s, qbs
