"""Micro-benchmark of the post-process kernels on config-5 style maps.  usage: python tools/bench_ccl.py [n_maps] [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops
from agenda_b200.synthetic import synthetic_heatmaps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
base = torch.from_numpy(synthetic_heatmaps(64, size, seed=0)).cuda()
maps = base.repeat((n + 63) // 64, 1, 1)[:n].contiguous()
for want_labels in (True, False):
    for _ in range(2):
        ops.ccl_bbox(maps, 0.5, 64, want_labels=want_labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.ccl_bbox(maps, 0.5, 64, want_labels=want_labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byts = n * size * size * (8 if want_labels else 4)
    print(f"ccl_bbox n={n} {size}x{size} labels={want_labels}: {ms:.3f} ms  {byts / ms / 1e6:.1f} GB/s  {n / ms * 1e3:.0f} maps/s")
# heat -> u8 image (64x64 -> 112x112), 3 planes + stack
heat = torch.rand((8192, 3, 64, 64), device="cuda")
for _ in range(2):
    ops.heat_postprocess_stack(heat, 112)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.heat_postprocess_stack(heat, 112)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
byts = 8192 * (3 * 64 * 64 * 4 + 3 * 112 * 112 * 2 + 112 * 112)
print(f"heat_postprocess_stack n=8192: {ms:.3f} ms  {byts / ms / 1e6:.1f} GB/s")
acc = torch.zeros((256, 77, 64, 64), device="cuda")
m32 = torch.rand((256, 77, 32, 32), device="cuda")
for _ in range(2):
    ops.heat_upsample_accum(m32, acc)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    ops.heat_upsample_accum(m32, acc)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
byts = 256 * 77 * (32 * 32 * 4 + 64 * 64 * 8)
print(f"heat_upsample_accum 32->64 planes={256*77}: {ms:.3f} ms  {byts / ms / 1e6:.1f} GB/s")
