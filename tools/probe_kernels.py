"""A few launches of one HBM-bound kernel at bench.py's probe shape (for ncu).
usage: python tools/probe_kernels.py upsample|postprocess|cross_bwd|cross_bwd_tc|self_bwd|glue|ccl"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "upsample"
dev = "cuda"
if which == "upsample":
    acc = torch.zeros((256 * 77, 64, 64), device=dev)
    m = torch.rand((256 * 77, 32, 32), device=dev)
    for _ in range(3):
        ops.heat_upsample_accum(m, acc)
elif which == "postprocess":
    heat = torch.rand((8192, 3, 64, 64), device=dev)
    for _ in range(3):
        ops.heat_postprocess_stack(heat, 112)
elif which == "cross_bwd":
    B, N, H, d, T = 2, 4096, 8, 40, 3
    q = torch.randn(B, N, H * d, device=dev).bfloat16()
    k = torch.randn(B, 77, H * d, device=dev).bfloat16()
    v = torch.randn(B, 77, H * d, device=dev).bfloat16()
    go = torch.randn_like(q)
    gm = torch.randn(B, T, N, device=dev)
    for _ in range(3):
        ops.attn_cross_bwd(q, k, v, go, gm, H, [5, 6, 7], 0)
elif which == "self_bwd":
    B, N, H, d = 2, 4096, 8, 40
    q, k, v, go = (torch.randn(B, N, H * d, device=dev).bfloat16() for _ in range(4))
    out = ops.attn_self(q, k, v, H)
    for _ in range(3):
        ops.attn_self_bwd(q, k, v, out, go, H)
elif which == "cross_bwd_tc":
    B, N, H, d, T = 2, 4096, 8, 40, 3
    q = torch.randn(B, N, H * d, device=dev).bfloat16()
    k = torch.randn(B, 77, H * d, device=dev).bfloat16()
    v = torch.randn(B, 77, H * d, device=dev).bfloat16()
    go = torch.randn_like(q)
    gm = torch.randn(B, T, N, device=dev)
    for _ in range(3):
        ops.attn_cross_bwd(q, k, v, go, gm, H, [5, 6, 7], 0)
elif which == "glue":
    x = torch.randn(16, 320, 64, 64, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.ones(320, device=dev).bfloat16()
    t = torch.randn(16, 4096, 320, device=dev).bfloat16()
    gg = torch.randn(16, 4096, 2560, device=dev).bfloat16()
    for _ in range(3):
        ops.groupnorm_nhwc(x, w, w, 32, 1e-5, True)
        ops.layernorm(t, w, w, 1e-5)
        ops.geglu(gg)
elif which == "ccl":
    from agenda_b200.synthetic import synthetic_heatmaps
    base = torch.from_numpy(synthetic_heatmaps(64, 512, seed=0)).to(dev)
    maps = base.repeat(32, 1, 1).contiguous()
    for _ in range(3):
        ops.ccl_bbox(maps, 0.5, 64)
torch.cuda.synchronize()
print("ok", which)
