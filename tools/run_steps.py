"""Profiling helper: run a few denoising steps of the SD-1.5 attention stack eagerly (no CUDA graph) so that
`ncu` sees every launch.  usage: python tools/run_steps.py [n_steps] [n_images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200.pipeline import sd15_pipeline

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pipe = sd15_pipeline(tokens=(5, 6, 7), num_steps=n_steps, use_cuda_graph=False)
hs, ctx = pipe.make_inputs(n_img, seed=0)
out = pipe.run_device(hs, ctx)
torch.cuda.synchronize()
print("ok", out["counts"].tolist())
