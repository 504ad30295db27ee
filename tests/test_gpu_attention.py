"""GPU parity (through the C ABI) of the attention kernels and the drop-in processor against the oracle and the
golden vectors recorded from the reference's hook.py."""
import os

import numpy as np
import pytest
import torch

from oracle import hook_oracle as O

pytestmark = pytest.mark.gpu

# Stated tolerances (BASELINE.json north_star): attention outputs max-abs 1e-2 on the bf16 tensor-core path,
# heat maps max-abs 1e-4 vs the fp32 reference.  The fp32 CUDA-core path is held to 2e-5.
TOL_BF16_OUT = 1e-2
TOL_F32_OUT = 2e-5
TOL_HEAT = 1e-4


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from agenda_b200 import ops as _ops
    return _ops


def _qkv(B, N, M, H, d, seed, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, N, H * d, generator=g) * gain
    k = torch.randn(B, M, H * d, generator=g) * gain
    v = torch.randn(B, M, H * d, generator=g)
    return q, k, v


@pytest.mark.parametrize("B,N,H,d", [(2, 64, 2, 40), (1, 300, 3, 64), (2, 256, 2, 80), (1, 64, 2, 160), (1, 130, 1, 8)])
def test_self_attention_f32(ops, B, N, H, d):
    q, k, v = _qkv(B, N, N, H, d, seed=N + d, gain=1.5)
    ref, _ = O.attention_core(q, k, v, H)
    out = ops.attn_self(q.cuda(), k.cuda(), v.cuda(), H, precision="fp32").cpu()
    assert (out - ref).abs().max().item() < TOL_F32_OUT
    # bf16 storage, fp32 math: compare against the oracle on the same bf16-rounded inputs
    qb, kb, vb = (t.bfloat16() for t in (q, k, v))
    refb, _ = O.attention_core(qb.float(), kb.float(), vb.float(), H)
    outb = ops.attn_self(qb.cuda(), kb.cuda(), vb.cuda(), H, precision="fp32").float().cpu()
    assert (outb - refb).abs().max().item() < TOL_BF16_OUT


@pytest.mark.parametrize("B,N,H,d,T", [(2, 64, 2, 40, None), (4, 256, 8, 40, [1, 5, 76]), (2, 100, 3, 64, [0]),
                                        (2, 16, 1, 160, [3, 4]), (2, 1024, 8, 80, [7, 9, 11])])
@pytest.mark.parametrize("is_train", [False, True])
def test_cross_attention_heat_f32(ops, B, N, H, d, T, is_train):
    M = 77
    q, k, v = _qkv(B, N, M, H, d, seed=N * 3 + d, gain=1.5)
    ref, p = O.attention_core(q, k, v, H)
    side = int(np.sqrt(N))
    b_first = 0 if is_train else B // 2
    toks = list(range(M)) if T is None else T
    if side * side == N:
        ref_maps = O.unravel_attn(p, H, is_train)[:, toks].reshape(B - b_first, len(toks), N)
    else:  # non-square N: same reduction without the (h,w) view
        pp = p.reshape(B, H, N, M)[b_first:].mean(1).permute(0, 2, 1)[:, toks]
        ref_maps = pp
    maps = torch.full((B - b_first, len(toks), N), 7.0, device="cuda")
    out = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=False).cpu()
    assert (out - ref).abs().max().item() < TOL_F32_OUT
    assert (maps.cpu() - ref_maps).abs().max().item() < 1e-6
    # accumulate mode adds on top
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=True)
    assert (maps.cpu() - 2 * ref_maps).abs().max().item() < 2e-6


@pytest.mark.parametrize("B,N,H,d,T", [(2, 64, 2, 40, None), (4, 256, 8, 40, [1, 5, 76]), (2, 100, 3, 64, [0]),
                                        (2, 16, 1, 160, [3, 4]), (2, 1024, 8, 80, [7, 9, 11]), (2, 4096, 8, 40, [5, 6, 7]),
                                        (4, 576, 5, 64, [1, 2, 3, 4]), (2, 256, 8, 160, None), (2, 130, 2, 40, [76])])
@pytest.mark.parametrize("is_train", [False, True])
def test_cross_attention_heat_tensor_core(ops, B, N, H, d, T, is_train):
    """bf16 activations -> tcgen05 kernel.  Oracle runs in fp32 on the same bf16-rounded inputs; heat maps must hold
    the fp32 tolerance (1e-4) because bf16 x bf16 products are exact in the fp32 accumulators."""
    M = 77
    q, k, v = _qkv(B, N, M, H, d, seed=N * 5 + d, gain=1.5)
    q, k, v = q.bfloat16(), k.bfloat16(), (v * 0.25).bfloat16()
    ref, p = O.attention_core(q.float(), k.float(), v.float(), H)
    b_first = 0 if is_train else B // 2
    toks = list(range(M)) if T is None else T
    ref_maps = p.reshape(B, H, N, M)[b_first:].mean(1).permute(0, 2, 1)[:, toks]
    maps = torch.full((B - b_first, len(toks), N), 7.0, device="cuda")
    out = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=False).float().cpu()
    assert (out - ref).abs().max().item() < TOL_BF16_OUT
    assert (maps.cpu() - ref_maps).abs().max().item() < 2e-6  # well inside TOL_HEAT
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=True)
    assert (maps.cpu() - 2 * ref_maps).abs().max().item() < 4e-6
    # the fp32 CUDA-core kernel agrees on the same inputs
    maps32 = torch.empty_like(maps)
    out32 = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps32, T, b_first, force_f32_kernel=True).float().cpu()
    assert (out32 - ref).abs().max().item() < TOL_BF16_OUT
    assert (maps32.cpu() - ref_maps).abs().max().item() < 2e-6
    # no heat requested
    out_nh = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, None).float().cpu()
    assert torch.equal(out_nh, out)


@pytest.mark.parametrize("qt", [2, 3])
@pytest.mark.parametrize("B,N,H,d,T", [(4, 1024, 8, 40, [1, 5, 76]), (2, 576, 5, 64, [1, 2, 3, 4]), (2, 4096, 8, 40, [5, 6, 7]),
                                        (2, 300, 2, 40, [0]), (2, 128, 1, 40, [7, 9]), (2, 640, 3, 64, None)])
@pytest.mark.parametrize("is_train", [False, True])
def test_cross_attention_resident_kv_kernel(ops, monkeypatch, B, N, H, d, T, is_train, qt):
    """The resident-K/V form (attn_cross_sm100_res.cu: K/V of all heads in shared memory, qt query tiles per CTA, two
    softmax warpgroups on alternate steps), forced through AGENDA_XRES_QT; must agree with the oracle and, bit for bit
    on the heat maps' inputs, with the per-(tile, head) kernel."""
    M = 77
    q, k, v = _qkv(B, N, M, H, d, seed=N * 7 + d + qt, gain=1.5)
    q, k, v = q.bfloat16(), k.bfloat16(), (v * 0.25).bfloat16()
    ref, p = O.attention_core(q.float(), k.float(), v.float(), H)
    b_first = 0 if is_train else B // 2
    if T is None:  # no heat maps requested
        monkeypatch.setenv("AGENDA_XRES_QT", str(qt))
        out = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, None).float().cpu()
        assert (out - ref).abs().max().item() < TOL_BF16_OUT
        return
    ref_maps = p.reshape(B, H, N, M)[b_first:].mean(1).permute(0, 2, 1)[:, T]
    monkeypatch.setenv("AGENDA_XRES", "0")
    maps0 = torch.zeros((B - b_first, len(T), N), device="cuda")
    out0 = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps0, T, b_first, accumulate=False)
    monkeypatch.setenv("AGENDA_XRES", "1")
    monkeypatch.setenv("AGENDA_XRES_QT", str(qt))
    maps = torch.full((B - b_first, len(T), N), 7.0, device="cuda")
    out = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=False)
    assert (out.float().cpu() - ref).abs().max().item() < TOL_BF16_OUT
    assert (maps.cpu() - ref_maps).abs().max().item() < 2e-6
    assert torch.equal(out, out0)                       # same MMAs, same softmax arithmetic
    assert (maps - maps0).abs().max().item() < 1e-6     # head sums in a different order
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=True)
    assert (maps.cpu() - 2 * ref_maps).abs().max().item() < 4e-6
    # DAAM-style per-head planes through the same kernel
    heads = torch.zeros((B - b_first, H, len(T), N), device="cuda")
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, heads, T, b_first, accumulate=True, per_head=True)
    ref_heads = p.reshape(B, H, N, M)[b_first:].permute(0, 1, 3, 2)[:, :, T]
    assert (heads.cpu() - ref_heads).abs().max().item() < 2e-6


@pytest.mark.parametrize("split", [2, 4, 8])
@pytest.mark.parametrize("B,N,H,d,T", [(2, 256, 8, 160, [5, 6, 7]), (2, 64, 8, 160, [1]), (4, 130, 8, 40, [0, 76]),
                                        (2, 1024, 8, 80, [7, 9, 11]), (2, 200, 4, 64, [1, 2, 3, 4, 5, 6, 7, 8])])
@pytest.mark.parametrize("is_train", [False, True])
def test_cross_attention_head_split_clusters(ops, monkeypatch, B, N, H, d, T, is_train, split):
    """Small layers run as clusters of `split` CTAs along the head axis (H / split heads each) and finish the head sum of
    the heat map through the leader's shared memory.  Forced through AGENDA_XSPLIT; must agree with the oracle, give
    the same attention output bit for bit as the unsplit kernel, and support accumulate and per-head planes."""
    if H % split:
        pytest.skip("heads not divisible")
    M = 77
    q, k, v = _qkv(B, N, M, H, d, seed=N * 3 + d + split, gain=1.5)
    q, k, v = q.bfloat16(), k.bfloat16(), (v * 0.25).bfloat16()
    ref, p = O.attention_core(q.float(), k.float(), v.float(), H)
    b_first = 0 if is_train else B // 2
    ref_maps = p.reshape(B, H, N, M)[b_first:].mean(1).permute(0, 2, 1)[:, T]
    monkeypatch.setenv("AGENDA_XRES", "0")
    monkeypatch.setenv("AGENDA_XSPLIT", "1")
    maps0 = torch.zeros((B - b_first, len(T), N), device="cuda")
    out0 = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps0, T, b_first, accumulate=False)
    monkeypatch.setenv("AGENDA_XSPLIT", str(split))
    maps = torch.full((B - b_first, len(T), N), 7.0, device="cuda")
    out = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=False)
    assert (out.float().cpu() - ref).abs().max().item() < TOL_BF16_OUT
    assert torch.equal(out, out0)
    assert (maps.cpu() - ref_maps).abs().max().item() < 2e-6
    assert (maps - maps0).abs().max().item() < 1e-6     # head sums grouped differently
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps, T, b_first, accumulate=True)
    assert (maps.cpu() - 2 * ref_maps).abs().max().item() < 4e-6
    maps2 = torch.full_like(maps, 3.0)                  # deterministic: same grouping, same bits
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps2, T, b_first, accumulate=False)
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, maps2, T, b_first, accumulate=True)
    assert torch.equal(maps2, maps)
    heads = torch.zeros((B - b_first, H, len(T), N), device="cuda")
    ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, heads, T, b_first, accumulate=True, per_head=True)
    ref_heads = p.reshape(B, H, N, M)[b_first:].permute(0, 1, 3, 2)[:, :, T]
    assert (heads.cpu() - ref_heads).abs().max().item() < 2e-6
    out_nh = ops.attn_cross_heat(q.cuda(), k.cuda(), v.cuda(), H, None)
    assert torch.equal(out_nh, out0)


def _golden_modules(g, name, tag, ctx_dim):
    from agenda_b200.sd_attention import SDAttention
    heads = int(g[f"{name}_heads"])
    C = g[f"{name}_hs"].shape[-1]
    m = SDAttention(C, ctx_dim if tag == "cross" else None, heads, C // heads)
    with torch.no_grad():
        m.to_q.weight.copy_(torch.from_numpy(g[f"{name}_{tag}_wq"]))
        m.to_k.weight.copy_(torch.from_numpy(g[f"{name}_{tag}_wk"]))
        m.to_v.weight.copy_(torch.from_numpy(g[f"{name}_{tag}_wv"]))
        m.to_out[0].weight.copy_(torch.from_numpy(g[f"{name}_{tag}_wo"]))
        m.to_out[0].bias.copy_(torch.from_numpy(g[f"{name}_{tag}_bo"]))
    return m.cuda()


@pytest.mark.parametrize("mode", ["infer", "train"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_processor_matches_reference_hook_golden(ops, golden_dir, mode, precision):
    """The drop-in processor reproduces what the reference's UNetCrossAttentionHooker recorded (hook.py executed
    unmodified, oracle/gen_golden.py): per-call outputs, per-call maps and compute_global_heat_map()."""
    from agenda_b200 import UNetCrossAttentionHooker
    g = np.load(os.path.join(golden_dir, f"hook_call_{mode}.npz"))
    ctx = torch.from_numpy(g["ctx"]).cuda()
    proc = UNetCrossAttentionHooker(is_train=(mode == "train"), latent_hw=16, precision=precision, record_maps=True)
    with pytest.raises(RuntimeError, match="No heat maps found."):
        proc.compute_global_heat_map()
    torch.backends.cuda.matmul.allow_tf32 = False
    for name in g["layer_names"]:
        hs = torch.from_numpy(g[f"{name}_hs"]).cuda()
        a_self = _golden_modules(g, name, "self", ctx.shape[-1])
        a_cross = _golden_modules(g, name, "cross", ctx.shape[-1])
        a_self.set_processor(proc); a_cross.set_processor(proc)
        with torch.no_grad():
            o_self = a_self(hs).cpu().numpy()
            o_cross = a_cross(hs, encoder_hidden_states=ctx).cpu().numpy()
        tol_self = TOL_F32_OUT * 5 if precision == "fp32" else TOL_BF16_OUT * 3  # x |to_out| gain
        assert np.abs(o_self - g[f"{name}_self_out"]).max() < tol_self, name
        # precision="bf16": the cross-attention call runs on the tensor cores too (split-precision kernel, all 77 token
        # maps): outputs to the bf16 tolerance (P V with bf16 P and V), maps far inside the 1e-4 heat tolerance
        tol_cross = TOL_F32_OUT * 5 if precision == "fp32" else TOL_BF16_OUT * 3
        assert np.abs(o_cross - g[f"{name}_cross_out"]).max() < tol_cross, name
        got = proc.cross_attn_maps[-1].cpu().numpy()
        assert got.shape == g[f"{name}_maps"].shape
        assert np.abs(got - g[f"{name}_maps"]).max() < (1e-6 if precision == "fp32" else 3e-5), name
    heat = proc.compute_global_heat_map().cpu().numpy()
    assert np.abs(heat - g["global"]).max() < (1e-5 if precision == "fp32" else 3e-5)
    proc.clear()
    assert proc.num_maps == 0 and proc.cross_attn_maps == []


def test_processor_token_subset_and_fused_accumulate(ops, golden_dir):
    """tokens=[...] + direct accumulation from the attention epilogue (layer at latent resolution)."""
    from agenda_b200 import UNetCrossAttentionHooker
    g = np.load(os.path.join(golden_dir, "hook_call_infer.npz"))
    ctx = torch.from_numpy(g["ctx"]).cuda()
    toks = [2, 7, 40]
    proc = UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=toks, precision="fp32")
    for name in g["layer_names"]:
        a = _golden_modules(g, name, "cross", ctx.shape[-1]); a.set_processor(proc)
        with torch.no_grad():
            a(torch.from_numpy(g[f"{name}_hs"]).cuda(), encoder_hidden_states=ctx)
    heat = proc.compute_global_heat_map().cpu().numpy()
    assert heat.shape == (1, 3, 16, 16)
    assert np.abs(heat - g["global"][:, toks]).max() < 1e-5


def test_processor_fp16_pipeline(ops):
    """Pipelines loaded with torch_dtype=torch.float16: modules and activations are fp16.  The processor runs the bf16
    tensor-core kernels on them and hands fp16 back; outputs and heat maps stay within the bf16 tolerances of an fp32
    evaluation of the same fp16-rounded weights and inputs (oracle processor_call, hook.py:83-122)."""
    from agenda_b200 import UNetCrossAttentionHooker
    from agenda_b200.sd_attention import SDAttention
    torch.manual_seed(4)
    C, H, hw = 320, 8, 16
    a_self = SDAttention(C, None, H, C // H).cuda().half()
    a_cross = SDAttention(C, 768, H, C // H).cuda().half()
    x = torch.randn(2, hw * hw, C, device="cuda").half()
    ctx = torch.randn(2, 77, 768, device="cuda").half()
    proc = UNetCrossAttentionHooker(is_train=False, latent_hw=hw, tokens=[3, 9])
    with torch.no_grad():
        y = proc(a_self, x)
        z = proc(a_cross, x, ctx)
    assert y.dtype == torch.float16 and z.dtype == torch.float16
    w = lambda m: m.weight.detach().float().cpu()
    yr, _ = O.processor_call(x.float().cpu(), None, w(a_self.to_q), w(a_self.to_k), w(a_self.to_v), w(a_self.to_out[0]),
                             a_self.to_out[0].bias.detach().float().cpu(), H, False)
    zr, mr = O.processor_call(x.float().cpu(), ctx.float().cpu(), w(a_cross.to_q), w(a_cross.to_k), w(a_cross.to_v),
                              w(a_cross.to_out[0]), a_cross.to_out[0].bias.detach().float().cpu(), H, False)
    assert (y.float().cpu() - yr).abs().max().item() < TOL_BF16_OUT * 3
    assert (z.float().cpu() - zr).abs().max().item() < TOL_BF16_OUT * 3
    heat = proc.compute_global_heat_map().cpu()
    assert (heat - mr[:, [3, 9]]).abs().max().item() < 2e-3   # q/k rounded fp16 -> bf16 before the scores


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows", ["broadcast", "per_query"])
def test_processor_attention_mask(ops, dtype, rows):
    """hook.py:92,108: an attention mask goes through attn.prepare_attention_mask and is ADDED to the scaled logits
    (baddbmm).  The SD UNets pass none; when one comes the processor takes the exact fp32 kernel: self- and
    cross-attention outputs and the heat maps against the oracle with the same additive mask (padding keys at -1e4 as
    diffusers builds them, plus finite per-key offsets; ragged N / M)."""
    from agenda_b200 import UNetCrossAttentionHooker
    from agenda_b200.sd_attention import SDAttention
    torch.manual_seed(11)
    C, H, hw, M = 320, 8, 12, 53
    B, N = 4, hw * hw
    a_self = SDAttention(C, None, H, C // H).cuda().to(dtype)
    a_cross = SDAttention(C, 768, H, C // H).cuda().to(dtype)
    x = torch.randn(B, N, C, device="cuda").to(dtype)
    ctx = torch.randn(B, M, 768, device="cuda").to(dtype)
    nq = 1 if rows == "broadcast" else N
    m_cross = torch.randn(B, nq, M, device="cuda")
    m_cross[:, :, 40:] = -10000.0          # padding keys
    m_cross[3, :, 17] = -10000.0       # one real token of the second conditional image
    m_self = torch.randn(B, nq, N, device="cuda") * 2
    m_self[:, :, -9:] = -10000.0
    proc = UNetCrossAttentionHooker(is_train=False, latent_hw=hw, tokens=[2, 17, 39])
    with torch.no_grad():
        y = a_self.set_processor(proc) or a_self(x, attention_mask=m_self.to(dtype))
        z = a_cross.set_processor(proc) or a_cross(x, encoder_hidden_states=ctx, attention_mask=m_cross.to(dtype))
    w = lambda m: m.weight.detach().float().cpu()
    rep = lambda m: m.to(dtype).float().cpu().repeat_interleave(H, dim=0)
    yr, _ = O.processor_call(x.float().cpu(), None, w(a_self.to_q), w(a_self.to_k), w(a_self.to_v), w(a_self.to_out[0]),
                             a_self.to_out[0].bias.detach().float().cpu(), H, False, mask=rep(m_self))
    zr, mr = O.processor_call(x.float().cpu(), ctx.float().cpu(), w(a_cross.to_q), w(a_cross.to_k), w(a_cross.to_v),
                              w(a_cross.to_out[0]), a_cross.to_out[0].bias.detach().float().cpu(), H, False, mask=rep(m_cross))
    tol_out = 2e-4 if dtype == torch.float32 else TOL_BF16_OUT * 3
    assert (y.float().cpu() - yr).abs().max().item() < tol_out
    assert (z.float().cpu() - zr).abs().max().item() < tol_out
    heat = proc.compute_global_heat_map().cpu()
    assert heat.shape == (B // 2, 3, hw, hw)
    assert (heat - mr[:, [2, 17, 39]]).abs().max().item() < (2e-6 if dtype == torch.float32 else 2e-3)
    # the masked token of batch element 3 (second image of the conditional half) carries no probability
    assert float(heat[1, 1].abs().max()) < 1e-6 and float(heat[0, 1].abs().max()) > 1e-4
