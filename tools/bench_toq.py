"""Micro-benchmark of the to_q forms feeding cross-attention: bf16-out GEMM (round 1), fp32-out GEMM, and the fp32-out
GEMM with the weight_lo correction (two GEMMs).  usage: python tools/bench_toq.py [images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from agenda_b200 import ops

images = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = 2 * images


def timed(fn, reps=10):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for N, C in [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]:
    M = B * N
    nbuf = max(1, int(300e6 // (M * C * 2)))
    xs = [torch.randn(M, C, device="cuda").bfloat16() for _ in range(nbuf)]
    w = (torch.randn(C, C, device="cuda") * C ** -0.5)
    w_hi = w.bfloat16()
    w_lo = (w - w_hi.float()).bfloat16()
    q32 = torch.empty(M, C, device="cuda")
    res = {}
    res["bf16-out"] = timed(lambda i: torch.nn.functional.linear(xs[i % nbuf], w_hi))
    res["fp32-out"] = timed(lambda i: torch.mm(xs[i % nbuf], w_hi.t(), out_dtype=torch.float32, out=q32))
    try:
        def two(i):
            torch.mm(xs[i % nbuf], w_hi.t(), out_dtype=torch.float32, out=q32)
            torch.addmm(q32, xs[i % nbuf], w_lo.t(), out_dtype=torch.float32, out=q32)
        res["fp32-out + addmm(lo) in place"] = timed(two)
        ref = (xs[0].float() @ (w_hi.float() + w_lo.float()).t())
        two(0)
        res["  max err vs fp32"] = float((q32 - ref).abs().max())
    except Exception as e:  # noqa: BLE001
        res["addmm in place"] = f"unsupported: {str(e)[:80]}"

    def three(i):
        q = torch.mm(xs[i % nbuf], w_hi.t(), out_dtype=torch.float32)
        q += torch.mm(xs[i % nbuf], w_lo.t(), out_dtype=torch.float32)
    res["mm + mm + add"] = timed(three)
    # K-concatenated single GEMM: [x | x] @ [w_hi ; w_lo]^T needs a duplicated activation matrix -> instead concatenate
    # along N and add the halves: one GEMM, 2x fp32 output
    res["agenda_linear_split_f32 (hi+lo)"] = timed(lambda i: ops.linear_split_f32(xs[i % nbuf], w_hi, w_lo))
    res["agenda_linear_split_f32 (hi)"] = timed(lambda i: ops.linear_split_f32(xs[i % nbuf], w_hi))
    pk = ops.linear_split_pack(w_hi, w_lo)
    res["packed weights (hi+lo)"] = timed(lambda i: ops.linear_split_f32_packed(xs[i % nbuf], pk))
    w_cat = torch.cat([w_hi, w_lo], 0)
    res["N-concat GEMM (2C fp32 out)"] = timed(lambda i: torch.mm(xs[i % nbuf], w_cat.t(), out_dtype=torch.float32))
    print(f"M={M} C={C}: " + " | ".join(f"{k}: {v:.1f}" if isinstance(v, float) else f"{k}: {v}" for k, v in res.items()), flush=True)
    del xs
