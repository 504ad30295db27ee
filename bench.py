#!/usr/bin/env python
"""bench.py — heat-map-labelled images/s through the AGenDA data-generation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--images-per-step 8]

One "step" = one batch of images through the whole hot path (BASELINE.json configs[1]): SD-1.5 shapes at a 64x64
latent, 50 denoising steps x 32 attention-processor calls (hook.py:83-122 semantics, CFG => UNet batch 2*images)
with per-token heat-map capture, hook.py:59-81 aggregation, data_generation.py:82-85 normalise/u8/resize,
postprocess_heatmap.py:44-46 stacking, and threshold/CCL/bbox.  The non-attention UNet layers
are out of scope and replaced by synthetic hidden states of the right shapes.

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = the same through the host-buffer API
(pinned host -> device copies of the step's inputs and device -> host reads of the results inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints
# "NCCL version ..." there when the box exports NCCL_DEBUG=VERSION), so fd 1 is pointed at stderr for the whole run
# and the result line goes out through a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "heatmap_labelled_images_per_s"
UNIT = "images/s"
NUM_DENOISE_STEPS = 50
TOKENS = (5, 6, 7)  # object word, fg token, bg token rows of the 77-token context


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-step", type=int, default=None, help="default 8 (sd15) / 16 (sd21)")
    ap.add_argument("--workload", default="sd15", choices=["sd15", "sd21", "config3", "train"],
                    help="sd15 = BASELINE configs[1] (the headline metric); sd21 = configs[3] (SD-2.1 768^2, 96^2 latent, "
                         "batch 16, 4 tokens) — informational; config3 = configs[2]: --num-images seeds sharded over the "
                         "ranks (strong scaling), final NCCL gather of heat maps + boxes, determinism check")
    ap.add_argument("--num-images", type=int, default=4096, help="config3: seeds 0..num_images-1 over all ranks")
    ap.add_argument("--denoise-steps", type=int, default=NUM_DENOISE_STEPS, help="config3: denoising steps per image")
    ap.add_argument("--dump", default=None, help="config3: rank 0 writes the gathered records to this .npz")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e through run_host (no upload / compute overlap)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-unet", action="store_true", help="skip the whole-UNet-step line (full_unet_step)")
    ap.add_argument("--ccl-maps", type=int, default=2048, help="512^2 maps for the post-process roofline probe")
    a = ap.parse_args()
    if a.images_per_step is None:
        a.images_per_step = 16 if a.workload == "sd21" else (2 if a.workload == "train" else 8)
    return a


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port of hook.py + numpy/PIL/scipy post-processing) on the host cores
# ------------------------------------------------------------------------------------------------------------

def cpu_reference_sample(sample_steps: int = 6, seed: int = 0):
    """Times `sample_steps` of the 50 denoising steps for ONE image (UNet batch 2) through the oracle port, plus the
    aggregation and post-processing once, and extrapolates to a full image.  Returns (images/s, detail dict)."""
    import numpy as np
    import torch
    from agenda_b200.sd_attention import AttentionStack, sd15_blocks
    from oracle import hook_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    blocks = sd15_blocks(64)
    stack = AttentionStack(blocks, 768, seed=0).float()
    hs, ctx = stack.make_inputs(2, "cpu", torch.float32, seed)
    maps = []
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(sample_steps):
            for b, a1, a2 in zip(blocks, stack.attn1, stack.attn2):
                x = hs[(b.hw, b.channels)]
                O.processor_call(x, None, a1.to_q.weight, a1.to_k.weight, a1.to_v.weight, a1.to_out[0].weight,
                                 a1.to_out[0].bias, b.heads, False)
                _, m = O.processor_call(x, ctx, a2.to_q.weight, a2.to_k.weight, a2.to_v.weight, a2.to_out[0].weight,
                                        a2.to_out[0].bias, b.heads, False)
                maps.append(m)
    t_attn = time.perf_counter() - t0
    t0 = time.perf_counter()
    heat = O.global_heat_map(maps, 64)[0]
    t_agg = (time.perf_counter() - t0) * (NUM_DENOISE_STEPS / sample_steps)
    t0 = time.perf_counter()
    planes = [O.heat_to_png_array(heat[t], 112) for t in TOKENS]
    O.stack_heatmaps(*planes)
    O.ccl_bbox(heat[TOKENS[0]], 0.5)
    t_post = time.perf_counter() - t0
    per_image = t_attn * (NUM_DENOISE_STEPS / sample_steps) + t_agg + t_post
    return 1.0 / per_image, {"cores": cores, "sample_s": t_attn + t_post,
                             "sample": f"{sample_steps} of {NUM_DENOISE_STEPS} denoising steps (32 attention calls "
                                       f"each, fp32, UNet batch 2 = 1 image) + aggregation + post-process, "
                                       f"extrapolated x{NUM_DENOISE_STEPS / sample_steps:g}"}


def heat_parity(pipe, hs_dev, ctx_dev, n_img, n_check):
    """max |heat - reference| over the first `n_check` images of the benchmarked batch.  Reference: the oracle port of
    hook.py in fp32 on the CPU with the pipeline's FP32 checkpoint weights and the same synthetic activations (cross
    attention only: the self-attention outputs do not feed the heat maps in this stack)."""
    import numpy as np
    import torch
    from oracle import hook_oracle as O
    pipe.use_cuda_graph = False
    pipe.num_steps = 1
    heat = pipe.run_device(hs_dev, ctx_dev)["heat"][:n_check].float().cpu().numpy()
    rows = list(range(n_check)) + [n_img + i for i in range(n_check)]       # [uncond..., cond...] of the checked images
    ref_stack = pipe.reference_stack()
    ctx = ctx_dev[rows].float().cpu()
    maps = []
    with torch.no_grad():
        for b, a2 in zip(pipe.blocks, ref_stack.attn2):
            x = hs_dev[(b.hw, b.channels)][rows].float().cpu()
            _, m = O.processor_call(x, ctx, a2.to_q.weight, a2.to_k.weight, a2.to_v.weight, a2.to_out[0].weight,
                                    a2.to_out[0].bias, b.heads, False)
            maps.append(m)
    ref = O.global_heat_map(maps, pipe.latent_hw)[:, pipe.tokens]
    n_px = n_diff = 0
    box_changed = 0
    for i in range(n_check):
        for t in range(3):
            a, r = O.heat_to_png_array(heat[i, t], pipe.image_size), O.heat_to_png_array(ref[i, t], pipe.image_size)
            n_px += a.size
            n_diff += int((a != r).sum())
        ba, br = O.ccl_bbox(heat[i, 0], pipe.thr)[1], O.ccl_bbox(ref[i, 0], pipe.thr)[1]
        box_changed += int(not (ba.shape == br.shape and np.array_equal(ba, br)))
    return {"value": float(np.abs(heat - ref).max()), "tolerance": 1e-4, "images_checked": n_check,
            "heat_max": float(ref.max()), "u8_pixels": n_px, "u8_pixels_differ": n_diff,
            "maps_with_box_change": box_changed,
            "against": "oracle (fp32 CPU restatement of hook.py:83-122,59-81) on the fp32 checkpoint weights, same "
                       "synthetic activations"}


# ncu `dram__bytes_read.sum + dram__bytes_write.sum` per launch of the probes below (profiles/r02_*_key_metrics.txt);
# None = not captured for that shape
NCU_TRAFFIC = {
    "heat_upsample_accum_32to64": 672.4e6,      # r02b_heat_upsample_persistent_full_key_metrics.txt: 403.7 MB read + 268.7 MB written
    "heat_postprocess_stack_64to112": 1067.7e6,  # r02b_postprocess_stack_persistent_full_key_metrics.txt: 402.8 MB + 664.9 MB
}


def _time_launches(fn, reps, graph=False):
    import torch
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if graph:   # kernels shorter than a ctypes call: replay the launches from a CUDA graph so the device is timed
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph:
        g.replay()
    else:
        for _ in range(reps):
            fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def probe_hbm_kernels(dev, hbm_gbs):
    """Standalone probes of the remaining HBM-bound kernels, inputs larger than L2: K3 (bicubic upsample + clamp +
    accumulate, hook.py:70-79), K4/K5 (normalise -> u8 -> PIL-bicubic resize -> stack, data_generation.py:82-85 +
    postprocess_heatmap.py:44-46) and K7 (cross-attention backward, training mode)."""
    import torch
    from agenda_b200 import ops
    out = {}
    # K3: 32 -> 64 on 19712 planes (0.7 GB of accumulator): read h*w*4 + read-modify-write L*L*8 per plane
    n_pl, h, L = 256 * 77, 32, 64
    acc = torch.zeros((n_pl, L, L), device=dev)
    m = torch.rand((n_pl, h, h), device=dev)
    ms = _time_launches(lambda: ops.heat_upsample_accum(m, acc), 5)
    by = n_pl * (h * h * 4 + L * L * 8)
    out["heat_upsample_accum_32to64"] = {"kernel": "heat_upsample_accum_tiled_kernel", "bound": "hbm", "achieved": by / ms / 1e6,
                                         "peak": hbm_gbs, "unit": "GB/s", "frac": by / ms / 1e6 / hbm_gbs, "ms": ms,
                                         "planes": n_pl, "traffic": NCU_TRAFFIC.get("heat_upsample_accum_32to64"),
                                         "note": "algorithmic bytes per plane = h*w*4 read + L*L*(4 read + 4 written)"}
    del acc, m
    # K4/K5: 8192 images, three 64x64 fp32 planes each -> planes u8 [3,112,112] + stack [112,112,3] + inverted bg [112,112]
    n_img, S = 8192, 112
    heat = torch.rand((n_img, 3, L, L), device=dev)
    ms = _time_launches(lambda: ops.heat_postprocess_stack(heat, S), 5)
    by = n_img * (3 * L * L * 4 + S * S * 7)
    out["heat_postprocess_stack_64to112"] = {"kernel": "postprocess_stack_kernel", "bound": "hbm", "achieved": by / ms / 1e6,
                                             "peak": hbm_gbs, "unit": "GB/s", "frac": by / ms / 1e6 / hbm_gbs, "ms": ms,
                                             "images": n_img, "traffic": NCU_TRAFFIC.get("heat_postprocess_stack_64to112"),
                                             "note": "algorithmic bytes per image = 3*L*L*4 read + S*S*(3+3+1) written; persistent CTAs "
                                                     "(coefficient tables once per CTA), maps held in registers, five-tap column-"
                                                     "owner passes; still bound by PIL's exact 8-bit fixed-point filter arithmetic "
                                                     "(issue slots 70 % busy), not by HBM"}
    del heat
    # K7: cross-attention backward at the 64x64 layer shape of one training sample pair (B = 2)
    B, N, H, d, T = 2, 4096, 8, 40, 3
    q = torch.randn(B, N, H * d, device=dev).bfloat16()
    k = torch.randn(B, 77, H * d, device=dev).bfloat16()
    v = torch.randn(B, 77, H * d, device=dev).bfloat16()
    go = torch.randn_like(q)
    gm = torch.randn(B, T, N, device=dev)
    toks = list(range(5, 5 + T))
    ms = _time_launches(lambda: ops.attn_cross_bwd(q, k, v, go, gm, H, toks, 0), 20, graph=True)
    by = 3 * B * N * H * d * 2 + 2 * B * 77 * H * d * 2 + 2 * B * 77 * H * d * 4 + gm.numel() * 4
    fl = 10.0 * B * H * N * 77 * d
    out["cross_attention_backward"] = {"kernel": "attn_cross_bwd_dq_kernel + attn_self_bwd_kernel<DKV, cross> (tcgen05)",
                                       "bound": "hbm", "achieved": by / ms / 1e6, "peak": hbm_gbs,
                                       "unit": "GB/s", "frac": by / ms / 1e6 / hbm_gbs, "ms": ms, "tflops": fl / ms / 1e9,
                                       "traffic": None,
                                       "note": "B=2, N=4096, d=40, T=3 (includes the memsets / casts of dK, dV): two tensor-core "
                                               "launches (dQ with the row softmax in registers; dK / dV on the transposed "
                                               "tiles, query range split over the SMs); 17 MB of operands, so launch- and "
                                               "latency-bound, not HBM-bound; the round-1 fp32 CUDA-core kernel took 198 us"}
    return out


def full_unet_step(dev, n_img, reps=2):
    """heat-map-labelled images/s when every denoising step runs the whole SD-1.x UNet skeleton (agenda_b200/unet.py:
    859.5 M random-init parameters, bf16, channels-last; attention through the processor and our kernels, everything
    else on cuDNN / cuBLAS), classifier-free guidance and the DDIM update — one CUDA graph per step, 50 steps."""
    import torch
    from agenda_b200.unet import UNetHeatmapPipeline
    up = UNetHeatmapPipeline(tokens=TOKENS, num_steps=NUM_DENOISE_STEPS, device=dev, cudnn_benchmark=True)
    lat, ctx = up.make_inputs(list(range(n_img)))
    up.run(lat, ctx)                                   # warm-up + graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = up.run(lat, ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": n_img / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "ms_per_denoise_step": ms / NUM_DENOISE_STEPS,
            "images_per_step": n_img, "denoise_steps": NUM_DENOISE_STEPS, "boxes_found": int(out["counts"].sum().item()),
            "note": "SD-1.x UNet skeleton (random init, 859.5 M parameters, bf16) at UNet batch %d with CFG 7.5 + DDIM; the "
                    "hidden states that reach the 32 attention calls are produced by the real dataflow and change every "
                    "step; convolutions / linears are library kernels (cuDNN / cuBLAS); GroupNorm(+SiLU), LayerNorm and GEGLU are "
                    "hand-written channels-last kernels (agenda_groupnorm_nhwc / agenda_layernorm / agenda_geglu)" % (2 * n_img)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup and 1):
        cpu_reference_sample(1)
    detail = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_reference_sample(6)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": detail["cores"], "kind": "port",
                             "sample": detail["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def workload_config(args, world):
    if args.workload == "sd21":
        return {"workload": "BASELINE configs[3]: SD-2.1-768 attention stack (16 transformer blocks, H=5/10/20/20, d=64, ctx "
                            "77x1024), 768^2 image / 96^2 latent, 50 denoising steps, batch "
                            f"{args.images_per_step} images per GPU (CFG => UNet batch {2 * args.images_per_step}), heat maps "
                            "for 4 tokens + u8 stacks + CCL boxes; non-attention UNet layers replaced by synthetic hidden "
                            "states; prompt K/V projected once per image batch",
                "images_per_step_per_gpu": args.images_per_step, "denoise_steps": NUM_DENOISE_STEPS, "tokens": 4,
                "parallelism": f"dp{world} (seed-sharded, final NCCL all_gather of boxes+heat maps)",
                "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"}
    return {"workload": "BASELINE configs[1]: SD-1.5 attention stack (16 transformer blocks, 32 processor calls per "
                        "UNet forward, H=8, d=40/80/160, ctx 77x768), 512^2 image / 64^2 latent, 50 denoising steps, "
                        f"batch {args.images_per_step} images per GPU (CFG => UNet batch {2 * args.images_per_step}), "
                        "heat maps for 3 tokens + u8 stacks (112^2) + CCL boxes; non-attention UNet layers replaced by "
                        "synthetic hidden states; to_k/to_v of the prompt embedding are projected once per image batch "
                        "(loop-invariant over the denoising steps), everything else runs at every step; bf16 "
                        "activations, fp32 checkpoint weights (bf16 copies + bf16 residuals for cross to_q/to_k), "
                        "cross-attention logits at fp32 accuracy (fp32-output to_q + split-precision tcgen05 kernel)",
            "images_per_step_per_gpu": args.images_per_step, "denoise_steps": NUM_DENOISE_STEPS, "tokens": len(TOKENS),
            "parallelism": f"dp{world} (seed-sharded, final NCCL all_gather of boxes+heat maps)",
            "l2": "per-step working set (hidden states 76 MB + Q/K/V/O activations > 300 MB) exceeds the 126 MB L2; "
                  "no explicit flush"}


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from agenda_b200 import _lib, ops
    from agenda_b200.pipeline import sd15_pipeline, sd21_pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_img = args.images_per_step

    sd21 = args.workload == "sd21"
    make_pipe = sd21_pipeline if sd21 else sd15_pipeline
    pipe = make_pipe(tokens=(4, 5, 6, 7) if sd21 else TOKENS, num_steps=NUM_DENOISE_STEPS, device=dev,
                     use_cuda_graph=not args.no_graph)
    big_n, big_d = (9216, 64) if sd21 else (4096, 40)
    hs_dev, ctx_dev = pipe.make_inputs(n_img, seed=rank)
    hs_host, ctx_host = pipe.make_inputs(n_img, seed=rank, pinned_host=True)
    staging = pipe.make_staging(hs_host, ctx_host)

    def gather(out):
        """final exchange: fixed-size records, one all_gather each."""
        if world == 1:
            return
        for name in ("counts", "boxes", "heat"):
            t = out[name].contiguous()
            buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(buf, t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_device():
        gather(pipe.run_device(hs_dev, ctx_dev))

    host_out = {}

    def step_host():
        out = pipe.run_host(hs_host, ctx_host, staging)
        host_out.update(out)
        gather(pipe.last_device_out)

    # ---- warm-up (also builds the CUDA graph) ----
    for _ in range(max(args.warmup, 1)):
        pipe.run_device(hs_dev, ctx_dev)
    barrier()

    # ---- timed: device-resident inputs ----
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launches
    ms_dev = timed(step_device, args.steps)
    launches = _lib.launches - l0
    graph_launch_note = None
    if not args.no_graph:
        # the 32 processor calls of a denoising step are replayed from a CUDA graph: count its kernel nodes of ours
        launches += (pipe._graph_launches * NUM_DENOISE_STEPS * args.steps) if hasattr(pipe, "_graph_launches") else 0
        graph_launch_note = "attention calls replayed from a CUDA graph"
    # ---- timed: end to end through the host-buffer API ----
    # run_host_batches: the public call for a stream of host batches; batch i+1's upload overlaps batch i's compute (every
    # batch's H2D and D2H are inside the timed region; only the first upload is exposed).  --e2e-serial times run_host
    # (upload, compute, download strictly in turn) instead.
    for _ in range(1):
        pipe.run_host(hs_host, ctx_host, staging)

    def e2e_stream():
        for out in pipe.run_host_batches(((hs_host, ctx_host) for _ in range(args.steps)), staging):
            host_out.update(out)
            gather(pipe.last_device_out)

    if args.e2e_serial:
        ms_e2e = timed(step_host, args.steps)
    else:
        e2e_stream()
        ms_e2e = timed(e2e_stream, 1)
    # the host link alone: one batch's inputs, pinned host -> device
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for k, v in hs_host.items():
        staging["hs"][k].copy_(v, non_blocking=True)
    staging["ctx"].copy_(ctx_host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_ms = h0.elapsed_time(h1)
    clocks = sampler.stop() if sampler else None

    images = n_img * world * args.steps
    value = images / (ms_dev / 1000.0)
    e2e_value = images / (ms_e2e / 1000.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_gbs, tf_burst, tf_sust, peak_src = measured_peaks()

    # ---- roofline of the dominant kernel: tcgen05 self-attention at N=4096, d=40 ----
    # CUDA-graph replays cannot hold per-kernel events, so the same step is replayed eagerly once with CUDA events
    # around every launch of the kernel on the launching stream.
    pipe_eager = pipe
    pipe_eager.use_cuda_graph = False
    _lib.event_sink = {"agenda_attn_self_fwd": [], "agenda_attn_self_fwd_strided": [], "agenda_attn_cross_fwd_heat": [],
                       "agenda_attn_cross_fwd_heat_x3": [], "agenda_attn_cross_fwd_heat_x3_hm": [],
                       "agenda_linear_split_f32": [], "agenda_linear_split_f32_heads": []}
    pipe_eager.num_steps = 5
    pipe_eager.run_device(hs_dev, ctx_dev)
    torch.cuda.synchronize()
    sink = _lib.event_sink
    _lib.event_sink = None

    def summarize(records, pick):
        tot_ms, tot_work, n = 0.0, 0.0, 0
        for s, e, a in records:
            w = pick(a)
            if w is None:
                continue
            tot_ms += s.elapsed_time(e); tot_work += w; n += 1
        return tot_ms, tot_work, n

    # args of agenda_attn_self_fwd[_strided]: q,k,v,out,dtype,B,H,N,d,[ld,]scale,stream
    self_calls = sink["agenda_attn_self_fwd"] + sink["agenda_attn_self_fwd_strided"]
    ms_k, flops, n_k = summarize(self_calls,
                                 lambda a: 4.0 * a[5] * a[6] * a[7] * a[7] * a[8] if a[7] == big_n else None)
    ms_all, flops_all, _ = summarize(self_calls, lambda a: 4.0 * a[5] * a[6] * a[7] * a[7] * a[8])
    achieved = flops / (ms_k / 1000.0) / 1e12 if ms_k > 0 else 0.0
    step_ms_eager_share = None
    kname = ("attn_self_sm100_v2_kernel<64,4,2,128,1,true> (N=9216, B*H=%d)" % (2 * n_img * 5) if sd21 else
             "attn_self_sm100_v2_kernel<40,3,3,64,1,true,true> (N=4096, B*H=%d)" % (2 * n_img * 8))
    roofline = {"kernel": kname, "bound": "tensor",
                "achieved": achieved, "peak": tf_sust, "unit": "TFLOP/s", "frac": achieved / tf_sust,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at B*H=128 from the ncu --set full capture
                # summarised in profiles/r01_self_attn_d40_full_key_metrics.txt (194.6 MB read + 39.6 MB written)
                "traffic": 234.2e6 if (n_img == 8 and not sd21) else None, "peak_source": f"{peak_src} bf16_tflops_sustained",
                "avg_launch_ms": ms_k / max(n_k, 1), "launches_timed": n_k,
                "share_of_step": (ms_k / 5.0 * NUM_DENOISE_STEPS) / (ms_dev / args.steps),
                "how": "CUDA events around each launch in an eager replay of 5 denoising steps; useful FLOPs "
                       "4*B*N*N*C with unpadded d=%d" % big_d}

    # ---- HBM-bound kernels: CCL on 512^2 maps (config 5 shape), standalone probe inside this run ----
    from agenda_b200.synthetic import synthetic_heatmaps
    n_maps = args.ccl_maps
    base = torch.from_numpy(synthetic_heatmaps(64, 512, seed=0)).to(dev)       # config-5 generator, 64 distinct maps
    maps = base.repeat((n_maps + 63) // 64, 1, 1)[:n_maps].contiguous()         # 2 GiB in + 2 GiB labels out >> L2
    del base
    for _ in range(2):
        ops.ccl_bbox(maps, 0.5, 64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        ops.ccl_bbox(maps, 0.5, 64)
    e1.record()
    torch.cuda.synchronize()
    ccl_ms = e0.elapsed_time(e1) / reps
    ccl_gbs = n_maps * 512 * 512 * 8 / (ccl_ms / 1000.0) / 1e9
    del maps

    extra = {"self_attention_all_layers": {"achieved_tflops": flops_all / (ms_all / 1000.0) / 1e12 if ms_all else 0,
                                           "ms_per_denoise_step": ms_all / 5.0},
             "ccl_bbox_512": {"bound": "hbm", "achieved": ccl_gbs, "peak": hbm_gbs, "unit": "GB/s",
                              "frac": ccl_gbs / hbm_gbs, "maps": n_maps, "ms": ccl_ms,
                              "traffic": 2.144e6,
                              "note": "BASELINE configs[4] generator (Gaussian blobs + noise floor), 64 distinct maps "
                                      "tiled, labels + boxes written; algorithmic bytes = H*W*(4 read + 4 written) per map; "
                                      "traffic = ncu dram bytes per map (profiles/r02b_ccl_bbox_cta_range_table_full_key_metrics.txt: "
                                      "a per-word range table lets the second pass re-read only the ~3 % of words that "
                                      "straddle the threshold)"}}

    # cross-attention + heat epilogue (K2, HBM-bound): the same eager replay, in-pipeline cache state (Q was just
    # written by the to_q GEMM).  The shipped path is the split-precision kernel (fp32 Q in, bf16 O out):
    # args: q,kv_blob,out,out_dtype,B,H,N,M,d,scale,token_idx,T,b_first,per_head,maps,accumulate,stream
    def cross_bytes_x3(a):
        B_, H_, N_, M_, d_, T_, bf_ = a[4], a[5], a[6], a[7], a[8], a[11], a[12]
        out_b = 4.0 if a[3] == 0 else 2.0
        return B_ * N_ * H_ * d_ * (4.0 + out_b) + 3.0 * B_ * M_ * H_ * d_ * 2 + (B_ - bf_) * T_ * N_ * 4.0

    # plain bf16 kernel (cross_logits="bf16"): q,k,v,out,dtype,B,H,N,M,d,scale,token_idx,T,b_first,maps,accumulate,stream
    def cross_bytes(a):
        B_, H_, N_, M_, d_, T_, bf_ = a[5], a[6], a[7], a[8], a[9], a[12], a[13]
        return 2.0 * B_ * N_ * H_ * d_ * 2 + 2.0 * B_ * M_ * H_ * d_ * 2 + (B_ - bf_) * T_ * N_ * 4.0
    x3_calls = sink["agenda_attn_cross_fwd_heat_x3"] + sink["agenda_attn_cross_fwd_heat_x3_hm"]
    if x3_calls:
        cross_calls, cb, n_at, kern = x3_calls, cross_bytes_x3, 6, "attn_cross_sm100_x3_kernel"
        note = "Q fp32 in + O bf16 out + K_hi,K_lo,V in + selected-token heat planes out"
        if sink["agenda_attn_cross_fwd_heat_x3_hm"]:
            note += " (Q in the chunk-major layout of the to_q GEMM, fetched by bulk copies)"
    else:
        cross_calls, cb, n_at, kern = sink["agenda_attn_cross_fwd_heat"], cross_bytes, 7, "attn_cross_sm100_res_kernel"
        note = "Q in + O out + K,V in + selected-token heat planes out (bf16)"
    ms_x, bytes_x, n_x = summarize(cross_calls, lambda a: cb(a) if a[n_at] == big_n else None)
    ms_x_all, _, _ = summarize(cross_calls, cb)
    if n_x:
        gbs = bytes_x / (ms_x / 1000.0) / 1e9
        extra["cross_attention_heat"] = {"kernel": kern, "bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                                         "frac": gbs / hbm_gbs, "avg_launch_ms": ms_x / n_x, "launches_timed": n_x,
                                         "ms_per_denoise_step_all_layers": ms_x_all / 5.0,
                                         "share_of_step": (ms_x_all / 5.0 * NUM_DENOISE_STEPS) / (ms_dev / args.steps),
                                         # r02_cross_x3_d40_full_key_metrics.txt: 90.1 MB read + 8.3 MB written (part of
                                         # Q, written by the to_q GEMM just before, is served by the 126 MB L2)
                                         "traffic": 98.5e6 if (kern == "attn_cross_sm100_x3_kernel" and n_img == 8 and not sd21) else None,
                                         "note": "N=%d layers; algorithmic bytes = %s" % (big_n, note)}

    # the fp32-output to_q GEMM of the cross-attention (x, w_hi, w_lo, out, M, K, N, ...): launches at the 64x64 layers
    ls_calls = sink["agenda_linear_split_f32"] + sink["agenda_linear_split_f32_heads"]
    ms_l, by_l, n_l = summarize(ls_calls, lambda a: (a[4] * a[5] * 2.0 + a[4] * a[6] * 4.0 + 2.0 * a[5] * a[6] * 2)
                                if a[4] == 2 * n_img * big_n else None)
    ms_l_all, _, _ = summarize(ls_calls, lambda a: 1.0)
    if n_l:
        extra["to_q_linear_split"] = {"kernel": "linear_split_kernel", "bound": "hbm", "achieved": by_l / (ms_l / 1000.0) / 1e9,
                                      "peak": hbm_gbs, "unit": "GB/s", "frac": by_l / (ms_l / 1000.0) / 1e9 / hbm_gbs,
                                      "avg_launch_ms": ms_l / n_l, "launches_timed": n_l,
                                      "ms_per_denoise_step_all_layers": ms_l_all / 5.0, "traffic": None,
                                      "note": "N=%d layers; algorithmic bytes = x bf16 in + Q fp32 out + both weight halves; "
                                              "two MMAs per activation (hi + lo weights)" % big_n}
    extra.update(probe_hbm_kernels(dev, hbm_gbs))

    # ---- parity of THIS run's heat maps: images 0 and 1 of the benchmarked batch against the oracle's fp32 evaluation
    #      of hook.py on the fp32 checkpoint weights (the checker; never on the product path) ----
    parity = None
    if not args.no_cpu_baseline:
        parity = heat_parity(pipe, hs_dev, ctx_dev, n_img, min(2, n_img))

    # ---- the same metric with the REST of the denoising step around the attention calls (SURVEY.md §8 f N4): SD-1.x
    #      UNet skeleton (convolutions / linears / norms on cuDNN / cuBLAS), CFG, DDIM, one CUDA graph per step ----
    full_step = None
    h2d_b, d2h_b = pipe.h2d_bytes(hs_host, ctx_host), pipe.d2h_bytes(host_out)
    if not args.no_unet and not sd21:
        del pipe, pipe_eager, hs_dev, ctx_dev, staging
        torch.cuda.empty_cache()
        full_step = full_unet_step(dev, n_img)

    cpu = None
    if not args.no_cpu_baseline:
        v, d = cpu_reference_sample(6)
        cpu = {"value": v, "unit": UNIT, "cores": d["cores"], "kind": "port", "sample": d["sample"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_ms_per_step_alone": h2d_ms,
                    "api": "run_host (serial)" if args.e2e_serial else "run_host_batches (next upload overlaps compute)",
                    "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": extra,
            "cpu_baseline": cpu, "heat_max_abs_err": parity, "full_unet_step": full_step}
    if graph_launch_note:
        line["gpu_launches_note"] = graph_launch_note
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config3(args):
    """BASELINE configs[2]: `--num-images` (prompt embedding, seed) pairs, rank r takes the seeds i = r (mod W)
    (SURVEY.md §8e), batches of --images-per-step, one final gather of fixed-size records over NCCL
    (agenda_b200.sharding.gather_records).  Strong scaling: the job is the same at every N.  One "step" = the whole job."""
    import torch
    import torch.distributed as dist
    from agenda_b200 import _lib
    from agenda_b200.pipeline import sd15_pipeline
    from agenda_b200.sharding import gather_records, shard_seeds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bs, n_total = args.images_per_step, args.num_images
    pipe = sd15_pipeline(tokens=TOKENS, num_steps=args.denoise_steps, device=dev, use_cuda_graph=not args.no_graph)
    seeds = shard_seeds(n_total, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job():
        local_rec = pipe.run_seeds(seeds, bs, getattr(pipe, "_seed_staging", None))
        return gather_records(local_rec, seeds, n_total)          # NCCL all_gather, seed-major re-order

    for _ in range(max(args.warmup, 1)):                             # warm-up: graph capture + a few batches
        pipe.run_seeds(seeds[:2 * bs], bs, getattr(pipe, "_seed_staging", None))
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        merged = job()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = _lib.launches - l0
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        # determinism across rank / batch composition: recompute seeds that (mostly) ran on OTHER ranks, in batches of a
        # different composition, on this GPU, and compare bit for bit with the gathered records
        picks = sorted({(i * (n_total // 16 or 1) + i % max(world, 1)) % n_total for i in range(16)})
        again = pipe.run_seeds(picks, bs, getattr(pipe, "_seed_staging", None))
        same = all(torch.equal(again[k], merged[k][picks]) for k in ("heat", "stack", "counts", "boxes"))
        if args.dump:
            import numpy as np
            np.savez(args.dump, **{k: v.cpu().numpy() for k, v in merged.items()})
        n_boxes = int(merged["counts"].sum().item())
        line = {"metric": METRIC, "value": n_total * args.steps / (ms / 1000.0), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[2]: SD-1.5 attention stack, {args.denoise_steps} denoising steps, "
                                       f"{n_total} synthetic (prompt embedding, seed) pairs, rank r takes seeds i = r mod "
                                       f"{world}, batches of {bs} images (inputs drawn on the device from each image's own "
                                       "seed), heat maps for 3 tokens + u8 stacks + CCL boxes, final NCCL all_gather of the "
                                       "records re-ordered by seed on every rank",
                           "num_images": n_total, "images_per_step_per_gpu": bs, "denoise_steps": args.denoise_steps,
                           "parallelism": f"dp{world} (seed-sharded, final NCCL all_gather)",
                           "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"},
                "gpu_launches": launches, "clocks": clocks,
                "gathered": {"images": int(merged["heat"].shape[0]), "boxes": n_boxes,
                             "bytes": int(sum(v.numel() * v.element_size() for v in merged.values()))},
                "determinism_check": {"seeds_recomputed_on_rank0": len(picks), "bit_identical": bool(same),
                                      "what": "heat, u8 stack, counts, boxes of seeds spread over all ranks' shards, recomputed "
                                              "in differently composed batches on rank 0's GPU"}}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
        if not same:
            raise SystemExit("config3: gathered records differ from a recomputation on rank 0")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """Training-mode processor (SURVEY.md §8 f N3; finetune_sd_token.py:1043-1069,1089): forward + backward of the SD-1.5
    attention stack — per block self -> cross -> self, UNet weights frozen, the prompt embedding carries the graph, loss =
    output term + L1 on the aggregated heat map of three tokens.  One "step" = one such train step at --images-per-step
    samples (is_train=True: no CFG pair).  Metric: train samples/s; roofline: the self-attention backward at N = 4096."""
    import torch
    from agenda_b200 import UNetCrossAttentionHooker, _lib, ops
    from agenda_b200.sd_attention import AttentionStack, sd15_blocks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = args.images_per_step
    stack = AttentionStack(sd15_blocks(), 768, seed=0).to(dev).to(torch.bfloat16)
    for p_ in stack.parameters():
        p_.requires_grad_(False)
    hs, ctx0 = stack.make_inputs(B, dev, torch.bfloat16)
    tgt = torch.rand(B, len(TOKENS), 64, 64, device=dev)
    proc = UNetCrossAttentionHooker(is_train=True, latent_hw=64, tokens=list(TOKENS), precision="bf16")

    static_ctx = ctx0.clone().requires_grad_(True)

    def step(ctx=None):
        proc.clear()
        if ctx is None:
            ctx = ctx0.clone().requires_grad_(True)
        loss = 0.0
        for b, a1, a2 in zip(stack.blocks, stack.attn1, stack.attn2):
            x = hs[(b.hw, b.channels)]
            y = proc(a2, proc(a1, x) + x, ctx)
            y = proc(a1, y)                  # activations now carry the graph: the self-attention backward runs
            loss = loss + y.float().pow(2).mean()
        loss = loss + 50.0 * (proc.compute_global_heat_map() - tgt).abs().mean()
        loss.backward()
        return loss.detach()

    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    l0 = _lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.launches - l0
    # the same step (forward + loss + backward, gradient into a static prompt-embedding leaf) replayed from ONE CUDA graph:
    # at batch 2 the eager step is bound by the host (Python autograd, ~400 launches of ~20 us kernels), the graph shows
    # what the kernels take
    graph_ms, graph_err = None, None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    static_ctx.grad = None
                    step(static_ctx)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            static_ctx.grad = None
            g_train = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_train):
                static_loss = step(static_ctx)
            for _ in range(2):
                g_train.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                g_train.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_ms = e0.elapsed_time(e1) / args.steps
            grad_ok = bool(torch.isfinite(static_ctx.grad.float()).all()) and float(static_ctx.grad.float().abs().max()) > 0
            if not grad_ok or not bool(torch.isfinite(static_loss)):
                graph_ms, graph_err = None, "graph replay produced a non-finite loss or an empty gradient"
        except Exception as exc:  # capture is an optimisation of the measurement, not of the product path
            graph_ms, graph_err = None, f"{type(exc).__name__}: {exc}"[:200]
    clocks = sampler.stop()
    _, _, tf_sust, peak_src = measured_peaks()
    # self-attention backward alone at the 64x64 layer shape
    N, H, d = 4096, 8, 40
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v, go = (torch.randn(B, N, H * d, device=dev, generator=g).bfloat16() for _ in range(4))
    o = ops.attn_self(q, k, v, H)
    ms_b = _time_launches(lambda: ops.attn_self_bwd(q, k, v, o, go, H), 10)
    flops = 10.0 * B * H * N * N * d
    line = {"metric": "train_samples_per_s", "value": B / (ms / 1000.0), "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "training-mode processor: SD-1.5 attention stack, 48 processor calls forward + backward "
                                   f"(self -> cross -> self per block), batch {B}, frozen weights, prompt embedding with grad, "
                                   "output + heat-map L1 loss; forward = the inference kernels, backward = agenda_attn_self_bwd "
                                   "+ agenda_attn_cross_bwd_tc (both tcgen05)", "samples_per_step": B},
            "gpu_launches": launches, "clocks": clocks, "loss": float(loss),
            "cuda_graph": {"ms_per_step": graph_ms, "samples_per_s": (B / (graph_ms / 1000.0)) if graph_ms else None,
                           "error": graph_err,
                           "note": "the same train step (forward + loss + backward) captured once and replayed; `value` "
                                   "above is the eager step"},
            "roofline": {"kernel": "attn_self_bwd_kernel<40, LSE|DQ|DKV> (N=4096, B*H=%d)" % (B * H), "bound": "tensor",
                         "achieved": flops / ms_b / 1e9, "peak": tf_sust, "unit": "TFLOP/s", "frac": flops / ms_b / 1e9 / tf_sust,
                         "traffic": None, "peak_source": f"{peak_src} bf16_tflops_sustained", "avg_launch_ms": ms_b,
                         "how": "CUDA events around 10 calls (4 launches each: Delta, LSE, dQ, dK+dV); useful FLOPs "
                                "10*B*H*N*N*d (five GEMMs); the kernels execute 8 GEMM units (S recomputed three times, dP twice, "
                                "dK and dV sharing one of each)"}}
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "train":
        run_train(a)
    elif a.workload == "config3":
        run_config3(a)
    else:
        run_ours(a)
