"""GPU parity of agenda_linear_split_f32 (to_q with an fp32 result: one tcgen05 GEMM over bf16 activations and the
hi / lo halves of an fp32 weight) against an fp64 matmul of the same values."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from agenda_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("M,K,N", [(128, 320, 320), (4096, 320, 320), (1000, 640, 640), (333, 1280, 1280), (65536, 320, 320),
                                    (64, 1280, 1280), (20000, 64, 160)])
@pytest.mark.parametrize("with_lo", [True, False])
def test_linear_split_f32(ops, M, K, N, with_lo):
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = torch.randn(N, K, generator=g) * K ** -0.5
    w_hi, w_lo = ops.split_bf16(w.cuda())
    out = ops.linear_split_f32(x.cuda(), w_hi, w_lo if with_lo else None)
    assert out.dtype == torch.float32 and out.shape == (M, N)
    w_eff = (w_hi.double() + (w_lo.double() if with_lo else 0)).cpu()
    rows = torch.randperm(M, generator=g)[:min(M, 512)]
    ref = x[rows].double() @ w_eff.t()
    err = (out[rows.cuda()].double().cpu() - ref).abs().max().item()
    assert err < 2e-5 * ref.abs().max().item() + 1e-6, err          # fp32 accumulation of exact products
    if with_lo:   # and that IS the fp32 checkpoint's projection (to 2^-17)
        ref32 = x[rows].double() @ w.double().t()
        assert (out[rows.cuda()].double().cpu() - ref32).abs().max().item() < 3e-5 * ref32.abs().max().item() + 1e-6
    # batched leading dims, determinism
    x3 = x.cuda().view(1, M, K)
    out2 = ops.linear_split_f32(x3, w_hi, w_lo if with_lo else None)
    assert out2.shape == (1, M, N) and torch.equal(out2[0], out)


def test_linear_split_f32_argument_errors(ops):
    from agenda_b200 import _lib
    x = torch.zeros(128, 320, device="cuda").bfloat16()
    with pytest.raises(_lib.AgendaError):   # N not a multiple of 160
        ops.linear_split_f32(x, torch.zeros(128, 320, device="cuda").bfloat16())
    with pytest.raises(TypeError):
        ops.linear_split_f32(x.float(), torch.zeros(320, 320, device="cuda").bfloat16())
    assert not ops.linear_split_f32_supported(x, torch.zeros(128, 320, device="cuda").bfloat16())
    assert ops.linear_split_f32_supported(x, torch.zeros(320, 320, device="cuda").bfloat16())


@pytest.mark.parametrize("B,Nq,K,H,d", [(2, 4096, 320, 8, 40), (3, 100, 640, 8, 80), (16, 64, 1280, 8, 160), (2, 33, 320, 4, 40),
                                         (5, 256, 1280, 8, 160), (1, 130, 320, 8, 40)])
def test_linear_split_heads_layout_equals_row_major(ops, B, Nq, K, H, d):
    """agenda_linear_split_f32_heads: the same numbers as agenda_linear_split_f32, bit for bit, in the chunk-major layout
    [B][H][d/40][N][40] (tiles that straddle two batch elements, ragged M)."""
    g = torch.Generator().manual_seed(Nq + d)
    x = torch.randn(B, Nq, K, generator=g).bfloat16().cuda()
    w = torch.randn(H * d, K, generator=g) * K ** -0.5
    w_hi, w_lo = ops.split_bf16(w.cuda())
    rows = ops.linear_split_f32(x, w_hi, w_lo)
    chunks = ops.linear_split_f32_heads(x, w_hi, w_lo, H)
    assert chunks.shape == (B, Nq, H * d) and chunks.buf.numel() == B * Nq * H * d
    assert torch.equal(chunks.to_rows(), rows)
    assert torch.equal(ops.linear_split_f32_heads(x, w_hi, None, H).to_rows(), ops.linear_split_f32(x, w_hi, None))


@pytest.mark.parametrize("M,K,N", [(4096, 320, 320), (1000, 640, 640), (333, 1280, 1280), (65536, 320, 320), (64, 1280, 1280),
                                    (20000, 64, 160)])
@pytest.mark.parametrize("with_lo", [True, False])
def test_linear_split_packed_weights_equal_tensor_map_loads(ops, M, K, N, with_lo):
    """agenda_linear_split_f32_packed (weights pre-packed into the stage images, one bulk copy per K block) against
    agenda_linear_split_f32 (tensor-map loads): the same MMAs on the same shared-memory images, so bit for bit."""
    g = torch.Generator().manual_seed(M + K + 1)
    x = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = torch.randn(N, K, generator=g) * K ** -0.5
    w_hi, w_lo = ops.split_bf16(w.cuda())
    lo = w_lo if with_lo else None
    packed = ops.linear_split_pack(w_hi, lo)
    assert packed.blob.numel() == N * K * 2 * (2 if with_lo else 1)
    assert torch.equal(ops.linear_split_f32_packed(x, packed), ops.linear_split_f32(x, w_hi, lo))
