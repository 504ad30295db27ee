"""Top CUDA kernels of one eager denoising step of the UNet skeleton (torch.profiler).  usage: python tools/profile_unet.py [images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from agenda_b200.unet import UNetHeatmapPipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2 and sys.argv[2] == "bench":
    torch.backends.cudnn.benchmark = True
pipe = UNetHeatmapPipeline(num_steps=2, use_cuda_graph=False)
lat, ctx = pipe.make_inputs(list(range(n)))
pipe.run(lat, ctx)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pipe.run(lat, ctx)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=60))
