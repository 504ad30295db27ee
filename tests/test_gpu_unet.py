"""GPU: the UNet skeleton + DDIM step around the attention processor (SURVEY.md §8 f N4): the heat-map path inside a
real denoising dataflow — hidden states that change per layer and per step — with per-call (teacher-forced) parity of
every cross-attention map against the oracle's fp32 evaluation of hook.py on the fp32 checkpoint weights."""
import numpy as np
import pytest
import torch

from oracle import hook_oracle as O

pytestmark = pytest.mark.gpu
TOL_HEAT = 1e-4


@pytest.fixture(scope="module")
def cuda_ok():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return True


def _pipe(**kw):
    from agenda_b200.unet import UNetHeatmapPipeline
    return UNetHeatmapPipeline(tokens=[5, 6, 7], latent_hw=16, max_boxes=32, **kw)


def test_unet_topology(cuda_ok):
    from agenda_b200.unet import SDUNet
    net = SDUNet(seed=1)
    mods = net.attention_modules()
    assert len(mods) == 32 and sum(m.is_cross for m in mods) == 16          # 16 transformer blocks
    names = [m.block_name for m in mods if m.is_cross]
    assert names[:2] == ["down0.0.attn2", "down0.1.attn2"] and "mid.attn2" in names and names[-1] == "up3.2.attn2"
    widths = sorted({m.to_q.in_features for m in mods})
    assert widths == [320, 640, 1280] and all(m.heads == 8 for m in mods)


def test_unet_pipeline_graph_equals_eager_and_is_deterministic(cuda_ok):
    a = _pipe(num_steps=3, use_cuda_graph=True)
    lat, ctx = a.make_inputs([3, 8])
    out1 = a.run(lat, ctx)
    out2 = a.run(lat, ctx)                                   # graph replay
    assert out1["heat"].shape == (2, 3, 16, 16) and out1["latents"].shape == (2, 4, 16, 16)
    for k in ("heat", "stack", "counts", "boxes", "latents"):
        assert torch.equal(out1[k], out2[k]), k
    assert torch.isfinite(out1["latents"].float()).all() and not torch.equal(out1["latents"], lat)
    b = _pipe(num_steps=3, use_cuda_graph=False)
    out3 = b.run(lat, ctx)
    assert (out3["heat"] - out1["heat"]).abs().max().item() < 1e-6
    # an image's result does not depend on its batch neighbours — up to the bf16 rounding of the library kernels, whose
    # algorithm choice (cuDNN / cuBLAS) changes with the batch size
    lat1, ctx1 = a.make_inputs([8])
    single = b.run(lat1, ctx1)
    assert (single["heat"][0] - out3["heat"][1]).abs().max().item() < 2e-4
    # downstream of the heat map everything is byte / integer work: bit-exact given OUR heat map
    heat = out1["heat"].cpu().numpy()
    for i in range(2):
        planes = [O.heat_to_png_array(heat[i, t], 112) for t in range(3)]
        rs, _ = O.stack_heatmaps(*planes)
        assert np.array_equal(out1["stack"][i].cpu().numpy(), rs)
        _, rb = O.ccl_bbox(heat[i, 0], 0.5)
        assert out1["counts"][i].item() == len(rb)


def test_unet_cross_attention_calls_match_fp32_reference(cuda_ok):
    """Teacher-forced parity inside the real dataflow: the inputs of all 16 cross-attention calls of one denoising step
    are recorded; the oracle evaluates hook.py's call (fp32, the fp32 checkpoint's weights) on them; the maps the
    processor recorded and the aggregated heat map must agree to 1e-4."""
    from agenda_b200 import UNetCrossAttentionHooker
    from agenda_b200.unet import SDUNet
    pipe = _pipe(num_steps=1, use_cuda_graph=False)
    proc = UNetCrossAttentionHooker(is_train=False, latent_hw=16, tokens=[5, 6, 7], record_maps=True)
    pipe.proc = proc
    pipe.unet.set_attn_processor(proc)
    recorded = []
    hooks = []
    for m in pipe.unet.attention_modules():
        if m.is_cross:
            hooks.append(m.register_forward_pre_hook(
                lambda mod, args, kwargs: recorded.append((mod.block_name, args[0].detach().float().cpu(),
                                                           kwargs["encoder_hidden_states"].detach().float().cpu())),
                with_kwargs=True))
    lat, ctx = pipe.make_inputs([0, 1])
    out = pipe.run(lat, ctx)
    for h in hooks:
        h.remove()
    assert len(recorded) == 16 and len(proc.cross_attn_maps) == 16
    ref_net = SDUNet(seed=0)                                   # the fp32 checkpoint the pipeline was built from
    ref_mods = {m.block_name: m for m in ref_net.attention_modules() if m.is_cross}
    ref_maps = []
    worst = 0.0
    for (name, hs, ehs), got in zip(recorded, proc.cross_attn_maps):
        a2 = ref_mods[name]
        _, m = O.processor_call(hs, ehs, a2.to_q.weight, a2.to_k.weight, a2.to_v.weight, a2.to_out[0].weight,
                                a2.to_out[0].bias, a2.heads, False)
        ref_maps.append(m)
        err = float((got.cpu() - m[:, [5, 6, 7]]).abs().max())
        worst = max(worst, err)
        assert err < TOL_HEAT, (name, err)
    ref_heat = O.global_heat_map(ref_maps, 16)[:, [5, 6, 7]]
    err = float(np.abs(out["heat"].cpu().numpy() - ref_heat).max())
    print(f"unet per-call heat err {worst:.2e}, aggregated {err:.2e}")
    assert err < TOL_HEAT


@pytest.mark.parametrize("B,C,H,W,silu", [(2, 320, 64, 64, True), (2, 2560, 8, 8, True), (3, 960, 16, 16, False),
                                           (1, 1920, 16, 16, True), (2, 640, 5, 7, False), (1, 32, 4, 4, True)])
def test_groupnorm_nhwc_matches_torch_fp32(cuda_ok, B, C, H, W, silu):
    """agenda_groupnorm_nhwc against torch.nn.functional.group_norm in fp32 on the same bf16-rounded input (the plain
    PyTorch fp32 reference of this floating-point op): one bf16 rounding of the result, so 1e-2 relative + 1e-2 absolute."""
    from agenda_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(B, C, H, W, generator=g) * 2 + 0.5).bfloat16()
    w = (torch.rand(C, generator=g) + 0.5).bfloat16()
    b = torch.randn(C, generator=g).bfloat16()
    ref = torch.nn.functional.group_norm(x.float(), 32, w.float(), b.float(), 1e-5)
    if silu:
        ref = torch.nn.functional.silu(ref)
    xd = x.cuda().contiguous(memory_format=torch.channels_last)
    assert ops.groupnorm_nhwc_supported(xd, 32)
    y = ops.groupnorm_nhwc(xd, w.cuda(), b.cuda(), 32, 1e-5, silu)
    assert y.shape == xd.shape and y.stride() == xd.stride()          # stays channels-last
    err = (y.float().cpu() - ref).abs()
    assert float((err - 1e-2 * ref.abs()).max()) < 1e-2, float(err.max())
    assert torch.equal(y, ops.groupnorm_nhwc(xd, w.cuda(), b.cuda(), 32, 1e-5, silu))   # no atomics: bit-reproducible
    # the [B, HW, C] view of the same memory gives the same numbers
    y3 = ops.groupnorm_nhwc(xd.permute(0, 2, 3, 1).reshape(B, H * W, C), w.cuda(), b.cuda(), 32, 1e-5, silu)
    assert torch.equal(y3.reshape(B, H, W, C).permute(0, 3, 1, 2), y)
    assert not ops.groupnorm_nhwc_supported(x.cuda(), 32) or H * W == 1 or C == 1   # NCHW memory is declined


def test_geglu_matches_torch_fp32(cuda_ok):
    from agenda_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(2, 77, 2 * 1280, generator=g) * 1.5).bfloat16()
    a, gate = x.float().chunk(2, dim=-1)
    ref = a * torch.nn.functional.gelu(gate)
    y = ops.geglu(x.cuda())
    assert y.shape == (2, 77, 1280) and y.dtype == torch.bfloat16
    err = (y.float().cpu() - ref).abs()
    assert float((err - 8e-3 * ref.abs()).max()) < 1e-3, float(err.max())


@pytest.mark.parametrize("M,C", [(2 * 4096, 320), (300, 640), (77, 1280), (5, 2560), (9, 8)])
def test_layernorm_matches_torch_fp32(cuda_ok, M, C):
    from agenda_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 3 + 1).bfloat16()
    w = (torch.rand(C, generator=g) + 0.5).bfloat16()
    b = torch.randn(C, generator=g).bfloat16()
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w.float(), b.float(), 1e-5)
    y = ops.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5)
    err = (y.float().cpu() - ref).abs()
    assert float((err - 8e-3 * ref.abs()).max()) < 1e-2, float(err.max())
    y0 = ops.layernorm(x.cuda().reshape(1, M, C), None, None, 1e-5)
    ref0 = torch.nn.functional.layer_norm(x.float(), (C,), None, None, 1e-5)
    assert float((y0.float().cpu().reshape(M, C) - ref0).abs().max()) < 3e-2


def test_resnet_block_fast_path_matches_plain_torch(cuda_ok):
    """ResnetBlock's inference fast path (bias-free convolutions, GroupNorm pre-add, fused bias + skip add) against the
    same block evaluated by plain torch modules in fp32."""
    from agenda_b200 import ops
    from agenda_b200.unet import ResnetBlock
    torch.manual_seed(0)
    for cin, cout in ((64, 64), (96, 64)):
        blk = ResnetBlock(cin, cout, temb=128)
        for p in blk.parameters():
            torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, mean=0.5, std=0.2)
        x = torch.randn(2, cin, 16, 16)
        temb = torch.randn(2, 128)
        fast = blk.to(device="cuda", dtype=torch.bfloat16).to(memory_format=torch.channels_last)
        with torch.no_grad():
            y = fast(x.cuda().bfloat16().contiguous(memory_format=torch.channels_last), temb.cuda().bfloat16())
            ref_blk = ResnetBlock(cin, cout, temb=128)
            ref_blk.load_state_dict({k: v.float().cpu() for k, v in fast.state_dict().items()})
            with torch.enable_grad():      # grad mode selects the plain torch path of the module
                ref = ref_blk(x.bfloat16().float().requires_grad_(True), temb.bfloat16().float()).detach()
        assert y.dtype == torch.bfloat16 and y.shape == ref.shape
        err = (y.float().cpu() - ref).abs().max().item()
        assert err < 3e-2 * ref.abs().max().item() + 1e-2, err
    # the GroupNorm pre-add alone
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 8, 8, generator=g).bfloat16()
    add = torch.randn(2, 64, generator=g).bfloat16()
    ref = torch.nn.functional.group_norm(x.float() + add.float()[:, :, None, None], 32, None, None, 1e-5)
    y = ops.groupnorm_nhwc(x.cuda().contiguous(memory_format=torch.channels_last), None, None, 32, 1e-5, False, add.cuda())
    assert (y.float().cpu() - ref).abs().max().item() < 3e-2
