import sys, os
sys.path.insert(0, "/root/repo")
import torch
from agenda_b200 import ops
for n_img, h in [(64, 32), (256, 32), (512, 32), (256, 16), (256, 8)]:
    acc = torch.zeros((n_img, 77, 64, 64), device="cuda")
    m = torch.rand((n_img, 77, h, h), device="cuda")
    for _ in range(2): ops.heat_upsample_accum(m, acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.heat_upsample_accum(m, acc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = n_img * 77 * (h * h * 4 + 64 * 64 * 8)
    print(f"upsample {h}->64 planes={n_img*77}: {ms:.3f} ms {byts/ms/1e6:.0f} GB/s")
