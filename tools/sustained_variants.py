import os, sys
sys.path.insert(0, os.getcwd())
import torch
from agenda_b200 import _lib
B, N, H, d = 16, 4096, 8, 40
torch.manual_seed(0)
q, k, v = (torch.randn(B, N, H * d, device="cuda").bfloat16() for _ in range(3))
out = torch.empty_like(q)
st = torch.cuda.current_stream().cuda_stream
flops = 4.0 * B * H * N * N * d
for var in (63, 64, 62, 63, 64, 62):
    def go(n):
        for _ in range(n):
            _lib.call("agenda_attn_self_fwd_variant", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, N, d, float(d ** -0.5), var, st)
    go(200); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(3000); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3000
    print(f"variant {var}: {ms:.4f} ms sustained  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
