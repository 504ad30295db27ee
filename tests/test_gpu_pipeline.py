"""GPU: the reference-facing surfaces end to end — HeatmapPipeline (device + host-buffer API, CUDA graph), the
daam-style trace shim, the host-buffer post-processing API and the two CLI mirrors — against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import hook_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda_ok():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return True


def _small_blocks():
    from agenda_b200.sd_attention import BlockSpec
    return [BlockSpec("down0", 16, 320, 8), BlockSpec("down1", 8, 640, 8), BlockSpec("mid", 4, 1280, 8),
            BlockSpec("up3", 16, 320, 8)]


TOL_HEAT = 1e-4   # BASELINE.json north_star: heat maps max-abs 1e-4 against the fp32 reference


def _oracle_heat(pipe, hs, ctx, steps, toks):
    """hook.py semantics in fp32 on the CPU with the pipeline's FP32 checkpoint weights (not the bf16 copies the device
    holds) and the same (bf16-valued) synthetic activations."""
    maps = []
    ref = pipe.reference_stack()
    for b, a2 in zip(pipe.blocks, ref.attn2):
        x = hs[(b.hw, b.channels)].float().cpu()
        w = [t.detach().float().cpu() for t in (a2.to_q.weight, a2.to_k.weight, a2.to_v.weight, a2.to_out[0].weight,
                                                a2.to_out[0].bias)]
        _, m = O.processor_call(x, ctx.float().cpu(), *w, b.heads, is_train=False)
        maps.append(m)
    return O.global_heat_map(maps * steps, pipe.latent_hw)[:, toks]


def _downstream_flips(heat, ref, size=112, thr=0.5):
    """How far the byte / integer outputs move when OUR heat map replaces the reference's: u8 pixels that differ (and by
    how much), and whether the CCL boxes differ.  heat / ref [n, T, L, L]."""
    n_px = n_diff = max_step = box_maps = 0
    for i in range(heat.shape[0]):
        for t in range(heat.shape[1]):
            a, b = O.heat_to_png_array(heat[i, t], size), O.heat_to_png_array(ref[i, t], size)
            d = np.abs(a.astype(np.int32) - b.astype(np.int32))
            n_px += d.size; n_diff += int((d > 0).sum()); max_step = max(max_step, int(d.max()))
        (_, ba), (_, bb) = O.ccl_bbox(heat[i, 0], thr), O.ccl_bbox(ref[i, 0], thr)
        box_maps += int(not (ba.shape == bb.shape and np.array_equal(ba, bb)))
    return {"u8_pixels": n_px, "u8_differ": n_diff, "u8_max_step": max_step, "maps_with_box_change": box_maps}


@pytest.mark.parametrize("graph", [False, True])
def test_pipeline_matches_oracle(cuda_ok, graph):
    from agenda_b200.pipeline import HeatmapPipeline
    toks = [2, 5, 9]
    pipe = HeatmapPipeline(_small_blocks(), 768, tokens=toks, num_steps=3, latent_hw=16, image_size=112,
                           use_cuda_graph=graph, max_boxes=32)
    hs, ctx = pipe.make_inputs(2, seed=3)  # 2 images -> UNet batch 4
    out = pipe.run_device(hs, ctx)
    ref = _oracle_heat(pipe, hs, ctx, 3, toks)          # [2, 3, 16, 16]; fp32 checkpoint weights, fp32 math
    heat = out["heat"].cpu().numpy()
    assert heat.shape == ref.shape
    # the cross-attention logits keep fp32 accuracy on the tensor cores (fp32-output to_q with the weight_lo correction
    # + the split-precision kernel): the stated 1e-4 holds against the fp32 reference on the fp32 weights
    assert np.abs(heat - ref).max() < TOL_HEAT
    # everything downstream of the heat map is integer/byte work: bit-exact given OUR fp32 heat map
    for i in range(2):
        planes = [O.heat_to_png_array(heat[i, t], 112) for t in range(3)]
        rs, ri = O.stack_heatmaps(*planes)
        assert np.array_equal(out["planes"][i].cpu().numpy(), np.stack(planes))
        assert np.array_equal(out["stack"][i].cpu().numpy(), rs) and np.array_equal(out["inv"][i].cpu().numpy(), ri)
        rl, rb = O.ccl_bbox(heat[i, 0], 0.5)
        assert out["counts"][i].item() == len(rb)
        assert np.array_equal(out["boxes"][i, :min(len(rb), 32)].cpu().numpy(), rb[:32])
    # second run (graph replay) reproduces the first bit for bit
    out2 = pipe.run_device(hs, ctx)
    assert torch.equal(out2["heat"], out["heat"]) and torch.equal(out2["boxes"], out["boxes"])


def test_pipeline_sd21_config4_shapes(cuda_ok):
    """BASELINE configs[3] at its real layer shapes: SD-2.1 768^2 -> 96^2 latent (N = 9216 / 2304 / 576 / 144, d = 64,
    H = 5 / 10 / 20 / 20, context 77 x 1024), four vehicle-class tokens, one image, two denoising steps; heat maps against
    the oracle, everything downstream bit-exact on our heat map."""
    from agenda_b200.pipeline import sd21_pipeline
    toks = [4, 5, 6, 7]
    pipe = sd21_pipeline(tokens=toks, num_steps=2, use_cuda_graph=True, max_boxes=32)
    hs, ctx = pipe.make_inputs(1, seed=11)
    out = pipe.run_device(hs, ctx)
    heat = out["heat"].cpu().numpy()
    ref = _oracle_heat(pipe, hs, ctx, 2, toks)
    assert heat.shape == ref.shape == (1, 4, 96, 96)
    assert np.abs(heat - ref).max() < TOL_HEAT
    planes = [O.heat_to_png_array(heat[0, t], 112) for t in range(3)]
    rs, ri = O.stack_heatmaps(*planes)
    assert np.array_equal(out["stack"][0].cpu().numpy(), rs) and np.array_equal(out["inv"][0].cpu().numpy(), ri)
    rl, rb = O.ccl_bbox(heat[0, 0], 0.5)
    assert out["counts"][0].item() == len(rb)
    assert np.array_equal(out["boxes"][0, :min(len(rb), 32)].cpu().numpy(), rb[:32])


def test_pipeline_sd15_real_shapes_fp32_reference(cuda_ok):
    """BASELINE configs[1] at its real layer shapes (SD-1.5, 64^2 latent: N = 4096 / 1024 / 256 / 64, d = 40 / 80 / 160,
    16 blocks), the benchmarked bf16 pipeline, two images, two denoising steps: heat maps within 1e-4 max-abs of the
    oracle's fp32 evaluation of hook.py on the FP32 checkpoint weights; and the damage downstream is counted: u8 pixels
    that differ from the reference's, and maps whose boxes change."""
    from agenda_b200.pipeline import sd15_pipeline
    toks = [5, 6, 7]
    pipe = sd15_pipeline(tokens=toks, num_steps=2, use_cuda_graph=True, max_boxes=64)
    hs, ctx = pipe.make_inputs(2, seeds=[0, 1])
    out = pipe.run_device(hs, ctx)
    heat = out["heat"].cpu().numpy()
    ref = _oracle_heat(pipe, hs, ctx, 2, toks)
    assert heat.shape == ref.shape == (2, 3, 64, 64)
    err = float(np.abs(heat - ref).max())
    assert err < TOL_HEAT, err
    flips = _downstream_flips(heat, ref)
    print(f"sd15 real shapes: heat max-abs err {err:.2e} (max {ref.max():.3f}); downstream {flips}")
    # a u8 level changes only where the normalised value sits within ~1e-4 * 255 of an integer: at most one level, rarely
    assert flips["u8_max_step"] <= 1 and flips["u8_differ"] <= 0.02 * flips["u8_pixels"]
    # the plain-bf16 logits path on the same inputs misses the tolerance (what round 1 shipped)
    fast = sd15_pipeline(tokens=toks, num_steps=2, use_cuda_graph=False, max_boxes=64, cross_logits="bf16")
    err_fast = float(np.abs(fast.run_device(hs, ctx)["heat"].cpu().numpy() - ref).max())
    print(f"  cross_logits='bf16': heat max-abs err {err_fast:.2e}")
    assert err_fast > 2 * err


def test_pipeline_fp32_activations(cuda_ok):
    """An fp32 pipeline (the reference's configuration, data_generation.py:30-31): fp32 modules and activations.  Cross
    attention takes the modules' own fp32 projections into the split-precision tensor-core kernel, self-attention the
    bf16 tensor-core kernel; heat maps within 1e-4, outputs fp32."""
    from agenda_b200.pipeline import HeatmapPipeline
    toks = [2, 5, 9]
    pipe = HeatmapPipeline(_small_blocks(), 768, tokens=toks, num_steps=2, latent_hw=16, dtype=torch.float32,
                           use_cuda_graph=False, max_boxes=32)
    hs, ctx = pipe.make_inputs(2, seeds=[4, 9])
    assert ctx.dtype == torch.float32
    out = pipe.run_device(hs, ctx)
    ref = _oracle_heat(pipe, hs, ctx, 2, toks)
    assert np.abs(out["heat"].cpu().numpy() - ref).max() < TOL_HEAT
    with torch.no_grad():
        y = pipe.stack(hs, ctx)
    assert y.dtype == torch.float32


def test_pipeline_host_api_equals_device_api(cuda_ok):
    from agenda_b200.pipeline import HeatmapPipeline
    pipe = HeatmapPipeline(_small_blocks(), 768, tokens=[1, 2, 3], num_steps=2, latent_hw=16, use_cuda_graph=True)
    hs_h, ctx_h = pipe.make_inputs(2, seed=1, pinned_host=True)
    assert all(v.is_pinned() for v in hs_h.values()) and ctx_h.is_pinned()
    host = pipe.run_host(hs_h, ctx_h)
    dev = pipe.run_device({k: v.cuda() for k, v in hs_h.items()}, ctx_h.cuda())
    for k in ("heat", "stack", "inv", "counts", "boxes"):
        assert torch.equal(host[k], dev[k].cpu()), k
    assert pipe.h2d_bytes(hs_h, ctx_h) > 0 and pipe.d2h_bytes(host) > 0


def test_pipeline_host_batches_prefetch_equals_one_by_one(cuda_ok):
    """run_host_batches overlaps the next batch's upload with the current batch's compute; records must equal run_host's."""
    from agenda_b200.pipeline import HeatmapPipeline
    pipe = HeatmapPipeline(_small_blocks(), 768, tokens=[1, 2, 3], num_steps=2, latent_hw=16, use_cuda_graph=True)
    batches = [pipe.make_inputs(2, seed=s, pinned_host=True) for s in (1, 5, 9)]
    want = [{k: v.clone() for k, v in pipe.run_host(hs, ctx).items()} for hs, ctx in batches]
    staging = pipe.make_staging(*batches[0])
    n = 0
    for host, ref in zip(pipe.run_host_batches(iter(batches), staging), want):
        for k in ("heat", "stack", "inv", "counts", "boxes"):
            assert torch.equal(host[k], ref[k]), (n, k)
        n += 1
    assert n == 3
    assert not torch.equal(want[0]["heat"], want[1]["heat"])
    assert list(pipe.run_host_batches(iter([]), staging)) == []


def test_pipeline_graph_replay_follows_in_place_input_updates(cuda_ok):
    """The captured graph reads the cached prompt K/V: overwriting the prompt embedding (and hidden states) in place
    between runs must be picked up (refresh_context_kv), exactly as an eager run on fresh tensors would."""
    from agenda_b200.pipeline import HeatmapPipeline
    pipe = HeatmapPipeline(_small_blocks(), 768, tokens=[1, 2, 3], num_steps=2, latent_hw=16, use_cuda_graph=True)
    eager = HeatmapPipeline(_small_blocks(), 768, tokens=[1, 2, 3], num_steps=2, latent_hw=16, use_cuda_graph=False)
    hs, ctx = pipe.make_inputs(2, seed=1)
    out1 = {k: v.clone() for k, v in pipe.run_device(hs, ctx).items()}
    hs2, ctx2 = pipe.make_inputs(2, seed=7)
    for k in hs:
        hs[k].copy_(hs2[k])
    ctx.copy_(ctx2)                                   # same buffers, new content -> graph is replayed, not recaptured
    out2 = pipe.run_device(hs, ctx)
    ref2 = eager.run_device(hs2, ctx2)
    assert not torch.equal(out2["heat"], out1["heat"])
    assert torch.equal(out2["heat"], ref2["heat"]) and torch.equal(out2["boxes"], ref2["boxes"])


def test_processor_context_kv_cache_invalidation(cuda_ok):
    """Cached to_k/to_v of the prompt embedding: reused for the same unmodified tensor, rebuilt after an in-place edit
    of the embedding or of a weight, never shared between different tensors."""
    from agenda_b200 import UNetCrossAttentionHooker
    from agenda_b200.sd_attention import SDAttention
    torch.manual_seed(0)
    attn = SDAttention(320, 768, 8, 40).cuda().bfloat16()
    x = torch.randn(2, 256, 320, device="cuda").bfloat16()
    ctx = torch.randn(2, 77, 768, device="cuda").bfloat16()
    cached = UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=[3])
    plain = UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=[3])
    plain.cache_context_kv = False
    with torch.no_grad():
        for step in range(3):
            assert torch.equal(cached(attn, x, ctx), plain(attn, x, ctx))
        assert len(cached._ctx_kv) == 1
        ctx.mul_(0.5)                                  # in-place edit of the embedding
        assert torch.equal(cached(attn, x, ctx), plain(attn, x, ctx))
        attn.to_v.weight.add_(0.01)                    # in-place edit of a weight
        assert torch.equal(cached(attn, x, ctx), plain(attn, x, ctx))
        other = torch.randn(2, 77, 768, device="cuda").bfloat16()
        assert torch.equal(cached(attn, x, other), plain(attn, x, other))
        assert torch.equal(cached.compute_global_heat_map(), plain.compute_global_heat_map())
        cached.clear()
        assert len(cached._ctx_kv) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,N,H,d,M,T,b_first", [(2, 256, 8, 40, 77, [2, 5], 0), (2, 64, 8, 160, 77, None, 1),
                                                 (3, 100, 2, 64, 77, [7, 7, 0], 1), (1, 33, 1, 80, 50, [49], 0),
                                                 (2, 1024, 8, 80, 77, None, 0), (2, 40, 3, 24, 96, [95, 1], 0)])
def test_cross_attention_backward_kernel(cuda_ok, monkeypatch, dtype, B, N, H, d, M, T, b_first):
    """agenda_attn_cross_bwd (csrc/attn_cross_bwd.cu) against fp32 autograd of the oracle's attention + head-mean maps on
    the same (dtype-rounded) inputs: dq, dk, dv with gradients arriving through BOTH the output and the maps, ragged
    N, repeated tokens, all tokens, a dropped unconditional half.  Tolerance relative to the largest entry: 1e-5 for
    fp32 tensors (fp32 math, different summation order), 1e-2 for bf16 tensors (gradients rounded to bf16)."""
    from agenda_b200.autograd import CrossAttentionHeatFn
    from oracle import hook_oracle as O
    g = torch.Generator(device="cuda").manual_seed(N + d)
    C = H * d
    q0 = torch.randn(B, N, C, device="cuda", generator=g).to(dtype)
    k0 = torch.randn(B, M, C, device="cuda", generator=g).to(dtype)
    v0 = torch.randn(B, M, C, device="cuda", generator=g).to(dtype)
    go = torch.randn(B, N, C, device="cuda", generator=g).to(dtype)
    nt = M if T is None else len(T)
    gm = torch.randn(B - b_first, nt, N, device="cuda", generator=g)
    q, k, v = (t.clone().requires_grad_(True) for t in (q0, k0, v0))
    out, maps = CrossAttentionHeatFn.apply(q, k, v, H, d ** -0.5, T, b_first)
    torch.autograd.backward((out, maps), (go, gm))
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q0, k0, v0))
    ref, p = O.attention_core(qr, kr, vr, H)
    ref_maps = p.reshape(B, H, N, M)[b_first:].mean(1).permute(0, 2, 1)[:, list(range(M)) if T is None else T]
    torch.autograd.backward((ref, ref_maps), (go.float(), gm))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for a, b_, name in ((q.grad, qr.grad, "dq"), (k.grad, kr.grad, "dk"), (v.grad, vr.grad, "dv")):
        assert a.dtype == dtype
        assert (a.float() - b_).abs().max().item() < tol * b_.abs().max().item(), name
    # the library-kernel formulation of the same backward agrees too
    monkeypatch.setenv("AGENDA_TORCH_CROSS_BWD", "1")
    q2, k2, v2 = (t.clone().requires_grad_(True) for t in (q0, k0, v0))
    out2, maps2 = CrossAttentionHeatFn.apply(q2, k2, v2, H, d ** -0.5, T, b_first)
    torch.autograd.backward((out2, maps2), (go, gm))
    for a, b_ in ((q2.grad, qr.grad), (k2.grad, kr.grad), (v2.grad, vr.grad)):
        assert (a.float() - b_).abs().max().item() < tol * b_.abs().max().item()


def test_cross_attention_backward_fp16(cuda_ok):
    """fp16 training (accelerate mixed_precision=fp16, which finetune_sd_token.py supports): forward AND backward of the
    cross-attention Function take fp16 tensors and hand fp16 gradients back."""
    from agenda_b200.autograd import CrossAttentionHeatFn
    from oracle import hook_oracle as O
    B, N, H, d, M, T = 2, 256, 8, 40, 77, [2, 5]
    g = torch.Generator(device="cuda").manual_seed(5)
    q0, k0, v0, go = (torch.randn(B, n, H * d, device="cuda", generator=g).half() for n in (N, M, M, N))
    gm = torch.randn(B, len(T), N, device="cuda", generator=g)
    q, k, v = (t.clone().requires_grad_(True) for t in (q0, k0, v0))
    out, maps = CrossAttentionHeatFn.apply(q, k, v, H, d ** -0.5, T, 0)
    assert out.dtype == torch.float16
    torch.autograd.backward((out, maps), (go, gm))
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q0, k0, v0))
    ref, p = O.attention_core(qr, kr, vr, H)
    ref_maps = p.reshape(B, H, N, M).mean(1).permute(0, 2, 1)[:, T]
    torch.autograd.backward((ref, ref_maps), (go.float(), gm))
    for a, b_ in ((q.grad, qr.grad), (k.grad, kr.grad), (v.grad, vr.grad)):
        assert a.dtype == torch.float16
        assert (a.float() - b_).abs().max().item() < 1e-2 * b_.abs().max().item()


def test_processor_under_inference_mode_and_on_other_current_device(cuda_ok):
    """torch.inference_mode() tensors carry no version counter (the prompt K/V cache key must not touch it), and the ops
    make the tensors' device current for the call (the library launches on the calling thread's current device)."""
    from agenda_b200 import UNetCrossAttentionHooker, ops
    from agenda_b200.sd_attention import SDAttention
    torch.manual_seed(2)
    attn = SDAttention(320, 768, 8, 40).cuda().bfloat16()
    proc = UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=[3, 9])
    with torch.inference_mode():
        x = torch.randn(2, 256, 320, device="cuda").bfloat16()
        ctx = torch.randn(2, 77, 768, device="cuda").bfloat16()
        y1 = proc(attn, x, ctx)
        y2 = proc(attn, x, ctx)              # second call: served from the prompt K/V cache
        heat = proc.compute_global_heat_map()
    assert torch.equal(y1, y2) and heat.shape == (1, 2, 16, 16) and torch.isfinite(heat).all()
    if torch.cuda.device_count() >= 2:
        q = torch.randn(2, 256, 320, device="cuda:1").bfloat16()
        with torch.cuda.device(0):
            o = ops.attn_self(q, q, q, 8)
        assert o.device.index == 1 and torch.isfinite(o.float()).all()


def _train_modules(C, heads, ctx_dim, seed):
    from agenda_b200.sd_attention import SDAttention
    torch.manual_seed(seed)
    self_attn = SDAttention(C, None, heads, C // heads).cuda()
    cross_attn = SDAttention(C, ctx_dim, heads, C // heads).cuda()
    return self_attn, cross_attn


def _oracle_block(x, ctx, self_attn, cross_attn, heads, is_train):
    """Reference math (oracle/hook_oracle.py:processor_call = hook.py:83-122) with autograd, fp32."""
    from oracle import hook_oracle as O

    def w(m):
        return m.weight.detach().float()
    y, _ = O.processor_call(x, None, w(self_attn.to_q), w(self_attn.to_k), w(self_attn.to_v), w(self_attn.to_out[0]),
                            self_attn.to_out[0].bias.detach().float(), heads, is_train)
    z, maps = O.processor_call(y, ctx, w(cross_attn.to_q), w(cross_attn.to_k), w(cross_attn.to_v),
                               w(cross_attn.to_out[0]), cross_attn.to_out[0].bias.detach().float(), heads, is_train)
    return z, maps


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
@pytest.mark.parametrize("hw,C,heads,is_train,tokens", [(16, 320, 8, True, [2, 5]), (8, 640, 8, True, None),
                                                        (16, 320, 8, False, [7])])
def test_processor_training_gradients_match_reference(cuda_ok, precision, tol, hw, C, heads, is_train, tokens):
    """SURVEY.md §8 f N3 / finetune_sd_token.py:1043-1069: self-attention -> cross-attention under autograd, loss =
    output term + L1-style term on the aggregated heat map of the selected tokens.  Gradients w.r.t. the latent
    features and the prompt embedding must match the reference processor's (oracle, fp32 autograd).  Relative
    tolerance on the largest gradient entry: 2e-4 with the fp32 kernels, 3e-2 with the bf16 tensor-core forward."""
    from agenda_b200 import UNetCrossAttentionHooker
    from oracle import hook_oracle as O
    B, N, L = 2, hw * hw, 2 * hw
    self_attn, cross_attn = _train_modules(C, heads, 768, seed=hw + C)
    g = torch.Generator(device="cuda").manual_seed(3)
    x0 = torch.randn(B, N, C, device="cuda", generator=g)
    c0 = torch.randn(B, 77, 768, device="cuda", generator=g)
    tgt = torch.rand(B if is_train else B // 2, 77 if tokens is None else len(tokens), L, L, device="cuda", generator=g)

    def loss_of(out, heat):
        return (out.float() ** 2).mean() + 50.0 * (heat - tgt).abs().mean()

    # ours
    x, c = x0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    proc = UNetCrossAttentionHooker(is_train=is_train, latent_hw=L, tokens=tokens, precision=precision)
    y = proc(self_attn, x)
    z = proc(cross_attn, y, c)
    assert len(proc.cross_attn_maps) == 1 and proc.cross_attn_maps[0].requires_grad
    heat = proc.compute_global_heat_map()
    loss = loss_of(z, heat)
    loss.backward()
    # reference
    xr, cr = x0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    zr, maps_r = _oracle_block(xr, cr, self_attn, cross_attn, heads, is_train)
    if tokens is not None:
        maps_r = maps_r[:, tokens]
    heat_r = torch.nn.functional.interpolate(maps_r, size=(L, L), mode="bicubic").clamp(min=0)   # hook.py:72, one map
    loss_r = loss_of(zr, heat_r)
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) < tol * max(1.0, abs(loss_r.item()))
    for a, b, name in ((x.grad, xr.grad, "d latent"), (c.grad, cr.grad, "d prompt")):
        scale = b.abs().max().item()
        assert scale > 0
        assert (a.float() - b).abs().max().item() < tol * scale, name
    # weight gradients exist too (the reference fine-tunes with frozen UNet weights, but nothing here assumes it)
    assert cross_attn.to_k.weight.grad is not None and self_attn.to_q.weight.grad is not None
    proc.clear()
    assert proc.num_maps == 0 and len(proc.cross_attn_maps) == 0


def test_processor_training_frozen_unet_prompt_gradient(cuda_ok):
    """The reference's actual setting: every UNet weight frozen, only the prompt embedding carries a graph
    (finetune_sd_token.py:754-757).  The FIRST self-attention call then sees nothing that requires grad and takes the
    plain kernels; from the first cross-attention on the graph must exist, and d loss / d prompt must be non-zero."""
    from agenda_b200 import UNetCrossAttentionHooker
    self_attn, cross_attn = _train_modules(320, 8, 768, seed=5)
    for m in (self_attn, cross_attn):
        for p_ in m.parameters():
            p_.requires_grad_(False)
    x = torch.randn(2, 256, 320, device="cuda")
    c = torch.randn(2, 77, 768, device="cuda", requires_grad=True)
    proc = UNetCrossAttentionHooker(is_train=True, latent_hw=16, tokens=[4])
    y = proc(self_attn, x)
    assert not y.requires_grad
    z = proc(cross_attn, y, c)
    z2 = proc(self_attn, z)          # now the activations carry the graph: self-attention backward is needed
    z3 = proc(cross_attn, z2, c)
    heat = proc.compute_global_heat_map()
    assert heat.shape == (2, 1, 16, 16) and heat.requires_grad
    (z3.float().pow(2).mean() + heat.mean()).backward()
    assert c.grad is not None and c.grad.abs().max().item() > 0
    with torch.no_grad():
        assert UNetCrossAttentionHooker(is_train=True)(self_attn, x).shape == x.shape


def test_trace_shim_end_to_end(cuda_ok):
    """data_generation.py:57-77 call sequence against the shim."""
    from agenda_b200.sd_attention import AttentionStack
    from agenda_b200.trace import trace
    blocks = _small_blocks()
    stack = AttentionStack(blocks, 768, seed=1).cuda().bfloat16()

    class Pipe:
        unet = stack
        tokenizer = None

    hs, ctx = stack.make_inputs(2, "cuda", torch.bfloat16, seed=2)  # batch 1 image + CFG
    with trace(Pipe(), tokens=[4, 7], latent_hw=16) as trc:
        with torch.no_grad():
            for _ in range(2):
                stack(hs, ctx)
        heat = trc.compute_global_heat_map()
    hm = heat.compute_word_heat_map("cars", token_idx=[4, 7]).heatmap
    assert hm.shape == (16, 16) and hm.dtype == torch.float32
    one = heat.compute_word_heat_map("cars", token_idx=[7]).heatmap
    assert torch.equal(one, heat.heat_maps[1])
    assert all(m.processor is None for m in list(stack.attn1) + list(stack.attn2))  # restored
    # daam mode: 3 of the 4 cross-attention layers (mid excluded)
    from agenda_b200 import UNetCrossAttentionHooker
    default = UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=[4])
    stack.set_attn_processor(default)  # in daam mode every non-hooked module keeps the processor it had
    with trace(Pipe(), tokens=[4], latent_hw=16, mode="daam") as trc:
        with torch.no_grad():
            stack(hs, ctx)
        assert trc.hooker.num_maps == 3 and default.num_maps == 1
    assert all(m.processor is default for m in list(stack.attn1) + list(stack.attn2))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_trace_daam_mode_matches_daam_oracle(cuda_ok, precision):
    """mode="daam": per-(layer, head) sums over the steps at native resolution, bicubic + clamp per (layer, head),
    mean over all of them, mid block and factor-8 layers left out (oracle/hook_oracle.py:daam_global_heat_map)."""
    from agenda_b200.sd_attention import AttentionStack, BlockSpec
    from agenda_b200.trace import trace
    blocks = [BlockSpec("down0", 16, 320, 8), BlockSpec("down1", 8, 640, 8), BlockSpec("down2", 2, 1280, 8),
              BlockSpec("mid", 4, 1280, 8), BlockSpec("up3", 16, 320, 8)]
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    stack = AttentionStack(blocks, 768, seed=4).cuda().to(dt)

    class Pipe:
        unet = stack
        tokenizer = None

    toks = [3, 6, 9]
    steps = 2
    hs, ctx = stack.make_inputs(2, "cuda", dt, seed=5)
    from agenda_b200 import UNetCrossAttentionHooker
    stack.set_attn_processor(UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=toks, precision=precision))
    with trace(Pipe(), tokens=toks, latent_hw=16, mode="daam", precision=precision) as trc:
        with torch.no_grad():
            for _ in range(steps):
                stack(hs, ctx)
        heat = trc.compute_global_heat_map().heat_maps.cpu().numpy()      # [T, 16, 16]
    sums = []
    for b, a2 in zip(blocks, stack.attn2):
        if b.name == "mid" or 16 // b.hw == 8:
            continue
        x = hs[(b.hw, b.channels)].float().cpu()
        q = O.head_to_batch_dim(x @ a2.to_q.weight.detach().float().cpu().T, b.heads)
        k = O.head_to_batch_dim(ctx.float().cpu() @ a2.to_k.weight.detach().float().cpu().T, b.heads)
        p = O.attention_probs(q, k, b.dim_head ** -0.5)                   # [B*H, N, 77]
        cond = p[b.heads:]                                                # conditional half (batch index 1)
        m = cond[:, :, toks].permute(0, 2, 1).reshape(b.heads, len(toks), b.hw, b.hw).numpy()
        sums.append(m * np.float32(steps))
    ref = O.daam_global_heat_map(sums, 16)
    assert heat.shape == ref.shape == (3, 16, 16)
    tol = 2e-3 * ref.max() + 1e-5 if precision == "bf16" else 1e-4 * ref.max() + 1e-6
    assert np.abs(heat - ref).max() < tol


def test_postprocess_host_api(cuda_ok):
    from agenda_b200 import postprocess
    rng = np.random.default_rng(0)
    heat = rng.random((3, 64, 64), dtype=np.float32) ** 3
    png = postprocess.heatmap_to_png_array(heat, 112)
    for i in range(3):
        assert np.array_equal(png[i], O.heat_to_png_array(heat[i], 112))
    st, inv = postprocess.stack_heatmaps(png[0], png[1], png[2])
    rs, ri = O.stack_heatmaps(png[0], png[1], png[2])
    assert np.array_equal(st, rs) and np.array_equal(inv, ri)
    big = O.synthetic_heatmaps(2, 256, seed=4)
    labels, boxes = postprocess.ccl_bbox(big, 0.5)
    for i in range(2):
        rl, rb = O.ccl_bbox(big[i], 0.5)
        assert np.array_equal(labels[i], rl) and np.array_equal(boxes[i], rb)
    lab1, box1 = postprocess.ccl_bbox(big[0], 0.5)
    assert np.array_equal(lab1, O.ccl_bbox(big[0], 0.5)[0])


def test_cli_postprocess_heatmap(cuda_ok, tmp_path):
    """The mirror of postprocess_heatmap.py writes the same PNG payloads as the reference's numpy lines."""
    from PIL import Image
    from agenda_b200 import postprocess_heatmap
    rng = np.random.default_rng(1)
    dirs = {k: tmp_path / f"daam_{k}_heatmaps" for k in ("cars", "fg", "bg")}
    arrays = {}
    for k, d in dirs.items():
        d.mkdir()
        for seed in (0, 1, 10):
            a = rng.integers(0, 256, (112, 112), dtype=np.uint8)
            arrays[(k, seed)] = a
            Image.fromarray(a).save(d / f"{seed}.png")
    n = postprocess_heatmap.main(["--save-dir", str(tmp_path), "--object-heatmap-path", "daam_cars_heatmaps",
                                  "--fg-heatmap-path", "daam_fg_heatmaps", "--bg-heatmap-path", "daam_bg_heatmaps"])
    assert n == 3
    for seed in (0, 1, 10):
        rs, ri = O.stack_heatmaps(arrays[("cars", seed)], arrays[("fg", seed)], arrays[("bg", seed)])
        st = Image.open(tmp_path / "daam_stack_heatmaps" / f"{seed}.png")
        assert st.mode == "RGB" and np.array_equal(np.asarray(st), rs)
        assert np.array_equal(np.asarray(Image.open(tmp_path / "daam_inv_heatmaps" / f"{seed}.png")), ri)


def test_cli_data_generation_synthetic(cuda_ok, tmp_path):
    """--synthetic: heat-map PNGs per word and seed, plus the COCO json of the CCL boxes (fixed 42.36 px rule), read
    back and checked against the oracle's boxes on the heat maps of the same seeds (every image is a function of its own
    seed: the CLI ran batches of 2, the check runs one batch of 3)."""
    import json
    from PIL import Image
    from agenda_b200 import data_generation
    from agenda_b200.pipeline import sd15_pipeline
    n = data_generation.main(["--synthetic", "--save-dir", str(tmp_path), "--num-images", "3", "--batch-size", "2",
                              "--num-inference-steps", "1", "--word_token_heatmaps", "cars", "fg", "bg",
                              "--token-indices", "5", "6", "7", "--image-size", "112"])
    assert n == 3
    pipe = sd15_pipeline(tokens=[5, 6, 7], num_steps=1, use_cuda_graph=False)
    # same batch size as the CLI's (the last CLI batch was padded to 2): bit-identical heat maps, hence identical bytes
    heats = []
    for chunk in ([0, 1], [2, 2]):
        hs, ctx = pipe.make_inputs(2, seeds=chunk)
        heats.append(pipe.run_device(hs, ctx)["heat"].cpu().numpy())
    heat = np.concatenate([heats[0], heats[1][:1]], 0)
    for w, word in enumerate(("cars", "fg", "bg")):
        for seed in range(3):
            im = Image.open(tmp_path / f"daam_{word}_heatmaps" / f"{seed}.png")
            assert im.size == (112, 112) and im.mode == "L"
            assert np.array_equal(np.asarray(im), O.heat_to_png_array(heat[seed, w], 112))
    coco = json.load(open(tmp_path / data_generation.COCO_FILE))
    assert coco["categories"] == [{"id": 1, "name": "small"}]
    assert [im["id"] for im in coco["images"]] == [0, 1, 2]
    assert [im["file_name"] for im in coco["images"]] == ["0.png", "1.png", "2.png"]
    for seed in range(3):
        want = O.coco_boxes_from_heat(heat[seed, 0], 0.5, 112)[:pipe.max_boxes]   # (the pipeline keeps the first max_boxes)
        got = [tuple(a["bbox"]) for a in coco["annotations"] if a["image_id"] == seed]
        assert got == want, seed
    assert all(a["area"] == a["bbox"][2] * a["bbox"][3] and a["category_id"] == 1 for a in coco["annotations"])
    assert len(coco["annotations"]) > 0


class _FakeTokenizer:
    def tokenize(self, text):
        return text.split()


class _FakePipeline:
    """What generate_with_pipeline touches on a diffusers StableDiffusionPipeline: .unet (modules with set_processor),
    .tokenizer, and pipeline(prompt, num_inference_steps=, generator=).images — here the attention stack on synthetic
    hidden states drawn from the generator, and a noise image instead of a VAE decode."""

    def __init__(self, stack, black_seeds=()):
        self.unet = stack
        self.tokenizer = _FakeTokenizer()
        self.black = set(black_seeds)
        self.inputs = {}

    def __call__(self, prompt, num_inference_steps=20, generator=None):
        from PIL import Image
        seed = generator.initial_seed()
        dt = next(self.unet.parameters()).dtype
        hs, ctx = self.unet.make_inputs(2, "cuda", dt, seed=seed)
        self.inputs[seed] = (hs, ctx)
        with torch.no_grad():
            for _ in range(num_inference_steps):
                self.unet(hs, ctx)
        rgb = np.zeros((64, 64, 3), np.uint8) if seed in self.black else \
            np.random.default_rng(seed).integers(1, 255, (64, 64, 3), dtype=np.uint8)

        class R:
            images = [Image.fromarray(rgb)]
        return R()


@pytest.mark.parametrize("mode", ["daam", "hook"])
def test_generation_loop_on_a_pipeline_object(cuda_ok, tmp_path, mode):
    """The non-synthetic branch of the data_generation.py mirror (generate_with_pipeline == data_generation.py:54-86),
    on a stand-in pipeline object: the reference's defaults (daam aggregation, precision from the pipeline dtype = fp32
    kernels for fp32 weights), word -> context rows through the tokenizer, black-image skip, PNG tree, COCO json."""
    import json
    from PIL import Image
    from agenda_b200 import UNetCrossAttentionHooker, data_generation
    from agenda_b200.sd_attention import AttentionStack, BlockSpec
    blocks = [BlockSpec("down0", 16, 320, 8), BlockSpec("down1", 8, 640, 8), BlockSpec("down2", 2, 1280, 8),
              BlockSpec("mid", 4, 1280, 8), BlockSpec("up3", 16, 320, 8)]
    stack = AttentionStack(blocks, 768, seed=4).cuda()               # fp32 weights, as the reference loads them
    stack.unet_config_sample_size = 16
    stack.set_attn_processor(UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=[1], precision="fp32"))
    pipe = _FakePipeline(stack, black_seeds=[1])
    args = data_generation.parse_args(["--save-dir", str(tmp_path), "--num-images", "3", "--num-inference-steps", "2",
                                       "--prompt", "an aerial view image with {} cars in {} utah",
                                       "--word_token_heatmaps", "cars", "utah", "--trace-mode", mode])
    assert args.trace_mode == mode and args.precision == "auto"
    stack.config = type("Cfg", (), {"sample_size": 16})()
    saved = data_generation.generate_with_pipeline(pipe, args, used=["<v0>", "<v1>"], words=["cars", "utah", "<v0>"])
    assert saved == 2                                                  # seed 1 produced a black image: skipped (:61-62)
    assert sorted(os.listdir(tmp_path / "images")) == ["0.png", "2.png"]
    prompt = "an aerial view image with <v0> cars in <v1> utah".split()
    rows = {w: prompt.index(w) + 1 for w in ("cars", "utah", "<v0>")}  # BOS offset (dataset.py:93)
    for seed in (0, 2):
        hs, ctx = pipe.inputs[seed]
        if mode == "daam":
            sums = []
            for b, a2 in zip(blocks, stack.attn2):
                if b.name == "mid" or 16 // b.hw == 8:
                    continue
                x = hs[(b.hw, b.channels)].float().cpu()
                q = O.head_to_batch_dim(x @ a2.to_q.weight.detach().float().cpu().T, b.heads)
                k = O.head_to_batch_dim(ctx.float().cpu() @ a2.to_k.weight.detach().float().cpu().T, b.heads)
                p = O.attention_probs(q, k, b.dim_head ** -0.5)[b.heads:]
                toks = sorted(rows.values())
                sums.append(p[:, :, toks].permute(0, 2, 1).reshape(b.heads, len(toks), b.hw, b.hw).numpy() * np.float32(2))
            ref = O.daam_global_heat_map(sums, 16)                      # rows in sorted-token order
        else:
            maps = []
            for b, a2 in zip(blocks, stack.attn2):
                w = [t.detach().float().cpu() for t in (a2.to_q.weight, a2.to_k.weight, a2.to_v.weight,
                                                        a2.to_out[0].weight, a2.to_out[0].bias)]
                maps.append(O.processor_call(hs[(b.hw, b.channels)].float().cpu(), ctx.float().cpu(), *w, b.heads, False)[1])
            ref = O.global_heat_map(maps * 2, 16)[0, sorted(rows.values())]
        order = {t: i for i, t in enumerate(sorted(rows.values()))}
        for word, row in rows.items():
            got = np.asarray(Image.open(tmp_path / f"daam_{word}_heatmaps" / f"{seed}.png")).astype(np.int32)
            want = O.heat_to_png_array(ref[order[row]], 112).astype(np.int32)
            d = np.abs(got - want)
            assert d.max() <= 1 and (d > 0).mean() < 0.02, (word, seed, d.max(), (d > 0).mean())
    coco = json.load(open(tmp_path / data_generation.COCO_FILE))
    assert [im["id"] for im in coco["images"]] == [0, 2] and coco["categories"][0]["name"] == "small"


def test_run_seeds_is_independent_of_batch_composition(cuda_ok):
    """Sharded generation: an image's records depend on its seed only — not on its neighbours in the
    batch, its position in it, or the padding of a short last batch."""
    from agenda_b200.pipeline import HeatmapPipeline
    pipe = HeatmapPipeline(_small_blocks(), 768, tokens=[2, 5, 9], num_steps=2, latent_hw=16, use_cuda_graph=True,
                           max_boxes=16)
    a = pipe.run_seeds([0, 1, 2, 3, 4, 5, 6], 4)            # 4 + 3 (padded)
    b = pipe.run_seeds([6, 3, 11, 0, 5], 4)                 # other neighbours, other positions
    for k in ("heat", "stack", "counts", "boxes"):
        assert a[k].shape[0] == 7 and b[k].shape[0] == 5
        for ia, ib in ((6, 0), (3, 1), (0, 3), (5, 4)):
            assert torch.equal(a[k][ia], b[k][ib]), (k, ia, ib)
    assert not torch.equal(a["heat"][0], a["heat"][1])
    empty = pipe.run_seeds([], 4)
    assert empty["heat"].shape[0] == 0 and empty["boxes"].shape == (0, 16, 5)
