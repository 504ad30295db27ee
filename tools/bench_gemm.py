"""Library GEMMs of the processor (to_q / fused qkv / to_out) at the SD-1.5 bench shapes: default cuBLASLt heuristic
vs torch's TunableOp.  usage: python tools/bench_gemm.py [tune]"""
import sys
import time
import torch
import torch.nn.functional as F

tune = len(sys.argv) > 1 and sys.argv[1] == "tune"
if tune:
    import torch.cuda.tunable as tn
    tn.enable(True)
    tn.tuning_enable(True)
    tn.set_max_tuning_duration(30)
    tn.set_max_tuning_iterations(20)
    try:
        tn.write_file_on_exit(False)
    except Exception:
        pass
shapes = []  # (M, K, N, bias)
for tokens, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
    M = 16 * tokens
    shapes += [(M, C, 3 * C, False), (M, C, C, True), (M, C, C, False)]
torch.manual_seed(0)
tot = 0.0
t0 = time.time()
for (M, K, N, bias) in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    b = torch.randn(N, device="cuda").bfloat16() if bias else None
    xs = [x.clone() for _ in range(max(1, int(200e6 // (x.numel() * 2))))]
    for _ in range(3):
        F.linear(x, w, b)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20):
            F.linear(xs[i % len(xs)], w, b)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    flops = 2.0 * M * K * N
    byts = 2.0 * (M * K + N * K + M * N)
    tot += us
    print(f"M={M:6d} K={K:5d} N={N:5d} bias={int(bias)}: {us:7.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s")
print(f"sum {tot:.1f} us  (tune={tune}, setup {time.time() - t0:.1f} s)")
