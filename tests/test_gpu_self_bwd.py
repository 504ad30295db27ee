"""GPU parity of the tcgen05 self-attention backward (agenda_attn_self_bwd) against fp32 autograd of the oracle's
restatement of hook.py:104-115 on the same bf16-rounded inputs.  Gradients are bf16 tensors produced from bf16 P / dS
operands: held to 2e-2 of the gradient's scale (the tolerance the training-mode tests use for bf16 tensors)."""
import pytest
import torch

from oracle import hook_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from agenda_b200 import ops as _ops
    return _ops


def _reference(q, k, v, go, H):
    q, k, v = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    d = q.shape[-1] // H
    qh, kh, vh = (O.head_to_batch_dim(t, H) for t in (q, k, v))
    p = O.attention_probs(qh, kh, d ** -0.5)
    out = O.batch_to_head_dim(torch.bmm(p, vh), H)
    out.backward(go.float())
    return out.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("B,N,H,d", [(2, 256, 2, 40), (1, 300, 3, 64), (2, 1024, 4, 80), (2, 64, 8, 160), (1, 130, 2, 40),
                                      (1, 4096, 2, 40), (2, 256, 8, 160), (1, 576, 5, 64)])
def test_self_attention_backward(ops, B, N, H, d):
    g = torch.Generator().manual_seed(N * 13 + d)
    q = (torch.randn(B, N, H * d, generator=g) * 1.5).bfloat16()
    k = (torch.randn(B, N, H * d, generator=g) * 1.5).bfloat16()
    v = torch.randn(B, N, H * d, generator=g).bfloat16()
    go = torch.randn(B, N, H * d, generator=g).bfloat16()
    out_ref, dq_r, dk_r, dv_r = _reference(q, k, v, go, H)
    out = ops.attn_self(q.cuda(), k.cuda(), v.cuda(), H)
    assert (out.float().cpu() - out_ref).abs().max().item() < 3e-2
    dq, dk, dv = ops.attn_self_bwd(q.cuda(), k.cuda(), v.cuda(), out, go.cuda(), H)
    for name, got, ref in (("dq", dq, dq_r), ("dk", dk, dk_r), ("dv", dv, dv_r)):
        got = got.float().cpu()
        assert torch.isfinite(got).all(), name
        err = (got - ref).abs().max().item()
        assert err < 2e-2 * ref.abs().max().item() + 1e-3, (name, err, ref.abs().max().item())


def test_self_attention_backward_through_autograd_function(ops):
    """SelfAttentionFn: forward = the tcgen05 kernel, backward = agenda_attn_self_bwd (no library attention anywhere)."""
    from agenda_b200 import autograd as ag
    import inspect
    assert "scaled_dot_product_attention" not in inspect.getsource(ag)
    B, N, H, d = 2, 512, 8, 40
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, N, H * d, generator=g).bfloat16() for _ in range(3))
    go = torch.randn(B, N, H * d, generator=g).bfloat16()
    _, dq_r, dk_r, dv_r = _reference(q, k, v, go, H)
    qd, kd, vd = (t.cuda().requires_grad_(True) for t in (q, k, v))
    out = ag.SelfAttentionFn.apply(qd, kd, vd, H, d ** -0.5, "bf16")
    out.backward(go.cuda())
    for got, ref in ((qd.grad, dq_r), (kd.grad, dk_r), (vd.grad, dv_r)):
        assert (got.float().cpu() - ref).abs().max().item() < 2e-2 * ref.abs().max().item() + 1e-3
    # the exact-parity path (precision="fp32"): fp32 GEMM formulas, per head
    qf, kf, vf = (t.float().cuda().requires_grad_(True) for t in (q, k, v))
    out32 = ag.SelfAttentionFn.apply(qf, kf, vf, H, d ** -0.5, "fp32")
    out32.backward(go.float().cuda())
    for got, ref in ((qf.grad, dq_r), (kf.grad, dk_r), (vf.grad, dv_r)):
        assert (got.cpu() - ref).abs().max().item() < 1e-4 * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("B,N,H,d", [(2, 256, 2, 40), (1, 300, 3, 64), (2, 1024, 4, 80), (1, 4096, 2, 40), (2, 256, 8, 160),
                                      (1, 130, 2, 40)])
def test_forward_emitted_lse_matches_and_feeds_the_backward(ops, B, N, H, d):
    """agenda_attn_self_fwd_lse: same output as agenda_attn_self_fwd bit for bit, lse = log2 sum_j 2^(scale log2e s_ij)
    against fp32 torch, and agenda_attn_self_bwd_lse with it agrees with the backward that runs its own LSE pass to bf16
    rounding.  Tolerance of lse: the row sum is the one the forward normalises with — at d = 40 it is accumulated by the PV MMA
    from the bf16-rounded probabilities, i.e. up to 2^-8 relative on a row one key dominates = 5.6e-3 in log2 units."""
    g = torch.Generator().manual_seed(N + d)
    q = (torch.randn(B, N, H * d, generator=g) * 1.5).bfloat16().cuda()
    k = (torch.randn(B, N, H * d, generator=g) * 1.5).bfloat16().cuda()
    v = torch.randn(B, N, H * d, generator=g).bfloat16().cuda()
    go = torch.randn(B, N, H * d, generator=g).bfloat16().cuda()
    out, lse = ops.attn_self_with_lse(q, k, v, H)
    assert lse is not None and lse.shape == (B, H, N)
    assert torch.equal(out, ops.attn_self(q, k, v, H))
    qh, kh = (O.head_to_batch_dim(t.float().cpu(), H) for t in (q, k))
    s = torch.bmm(qh, kh.transpose(1, 2)) * (d ** -0.5) * 1.4426950408889634
    ref = torch.logsumexp(s * 0.6931471805599453, dim=-1) * 1.4426950408889634
    assert (lse.cpu().reshape(B * H, N) - ref).abs().max().item() < 8e-3
    a = ops.attn_self_bwd(q, k, v, out, go, H, lse=lse)
    b = ops.attn_self_bwd(q, k, v, out, go, H)
    for x, y in zip(a, b):
        assert (x.float() - y.float()).abs().max().item() <= 2e-2 * y.float().abs().max().item() + 1e-3
    small, none = ops.attn_self_with_lse(q[:, :64], k[:, :64], v[:, :64], H)
    assert none is None and small.shape == (B, 64, H * d)
