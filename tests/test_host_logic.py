"""CPU: host-side logic — seed sharding + gather (world_size-2 gloo), the fixed-size box rule, CLI flag parity with
the reference scripts, trace shim plumbing."""
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT_ not in sys.path:
    sys.path.insert(0, ROOT_)
from oracle import hook_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_seeds_partition():
    from agenda_b200.sharding import padded_count, shard_seeds
    for n, w in [(10, 1), (10, 2), (4096, 8), (7, 4), (3, 8), (0, 2)]:
        parts = [shard_seeds(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert all(all(s % w == r for s in p) for r, p in enumerate(parts))
        assert max(len(p) for p in parts) <= padded_count(n, w)
    with pytest.raises(ValueError):
        shard_seeds(4, 2, 2)


GLOO_WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from agenda_b200.sharding import shard_seeds, gather_records
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 7
    seeds = shard_seeds(n, rank, world)
    local = {"counts": torch.tensor([s * 10 for s in seeds], dtype=torch.int32),
             "boxes": torch.stack([torch.full((4, 5), s, dtype=torch.int32) for s in seeds]) if seeds else torch.zeros((0, 4, 5), dtype=torch.int32),
             "heat": torch.stack([torch.full((2, 3, 3), float(s)) for s in seeds]) if seeds else torch.zeros((0, 2, 3, 3))}
    out = gather_records(local, seeds, n)
    assert out["counts"].tolist() == [s * 10 for s in range(n)], out["counts"]
    assert all(int(out["boxes"][s, 0, 0]) == s for s in range(n))
    assert all(float(out["heat"][s, 1, 2, 2]) == s for s in range(n))
    assert out["heat"].shape == (n, 2, 3, 3)
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
    dist.destroy_process_group()
""")


def test_gather_records_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GLOO_OK" in res.stdout


def test_gather_records_single_process():
    from agenda_b200.sharding import gather_records, shard_seeds
    seeds = shard_seeds(5, 0, 1)
    out = gather_records({"x": torch.arange(5)}, seeds, 5)
    assert out["x"].tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        gather_records({"x": torch.arange(4)}, seeds, 5)


def test_fixed_size_boxes_matches_reference_rule():
    from agenda_b200.postprocess import fixed_size_boxes
    rng = np.random.default_rng(0)
    xywh = np.concatenate([rng.uniform(0, 100, (200, 2)), rng.uniform(1, 40, (200, 2))], 1)
    xywh[:, 2] = np.minimum(xywh[:, 2], 112 - xywh[:, 0])
    xywh[:, 3] = np.minimum(xywh[:, 3], 112 - xywh[:, 1])
    got = fixed_size_boxes(xywh)
    for i, (x, y, w, h) in enumerate(xywh):
        assert tuple(got[i]) == O.fixed_size_box_rule(x, y, x + w, y + h), i
    centre = fixed_size_boxes(np.array([[50, 50, 10, 10, 100]]))[0]
    assert np.allclose(centre, [55 - 21.18, 55 - 21.18, 42.36, 42.36])
    assert fixed_size_boxes(np.zeros((0, 5))).shape == (0, 4)
    assert fixed_size_boxes(np.array([1.0, 2.0, 3.0, 4.0])).shape == (1, 4)


def test_ccl_boxes_to_coco_json(tmp_path):
    """CCL output (counts, padded int boxes on the L x L map) -> image-grid fixed-size boxes -> COCO dict in the
    reference's layout (refine_label.py:28-48,126-135; [x, y, w, h] top-left, Data/README.md:7) -> json round trip."""
    import json
    from agenda_b200 import postprocess
    counts = np.array([2, 0, 1], dtype=np.int32)
    boxes = np.zeros((3, 4, 5), dtype=np.int32)
    boxes[0, 0] = (10, 12, 6, 4, 20); boxes[0, 1] = (0, 60, 3, 4, 9); boxes[2, 0] = (60, 1, 4, 2, 8)
    per = postprocess.ccl_boxes_to_coco_boxes(counts, boxes, map_size=64, image_size=112)
    assert [p.shape for p in per] == [(2, 4), (0, 4), (1, 4)]
    s = 112 / 64
    for got, raw in ((per[0][0], boxes[0, 0]), (per[0][1], boxes[0, 1]), (per[2][0], boxes[2, 0])):
        x, y, w, h = (float(v) * s for v in raw[:4])
        assert tuple(got) == O.fixed_size_box_rule(x, y, x + w, y + h)
    coco = postprocess.coco_annotations(["7.png", "8.png", "9.png"], per, (112, 112), image_ids=[7, 8, 9])
    assert coco["categories"] == [{"id": 1, "name": "small"}]
    assert [im["id"] for im in coco["images"]] == [7, 8, 9] and coco["images"][0]["width"] == 112
    assert [a["image_id"] for a in coco["annotations"]] == [7, 7, 9]
    a0 = coco["annotations"][0]
    assert a0["iscrowd"] == 0 and a0["category_id"] == 1 and a0["area"] == a0["bbox"][2] * a0["bbox"][3]
    path = tmp_path / "ann.json"
    postprocess.write_coco_json(str(path), coco)
    assert json.load(open(path)) == coco


def _flags(src):
    return sorted(set(re.findall(r'add_argument\(\s*"(--[A-Za-z_\-]+)"', src)))


@pytest.mark.parametrize("name", ["data_generation.py", "postprocess_heatmap.py"])
def test_cli_flags_cover_reference(name):
    """Every flag of the reference CLI exists, with the same default, in the mirror.  The reference's flag list is
    recorded here (data_generation.py:11-23, postprocess_heatmap.py:8-17) because /root/reference is absent on the
    GPU box; when it is present the recorded list is cross-checked against the file itself."""
    recorded = {
        "data_generation.py": ["--image-size", "--initialize_token", "--learnable-tokens-embedding-path", "--num-images",
                               "--pretrained-model-path", "--prompt", "--save-dir", "--store_learnable_token_heatmaps",
                               "--word_token_heatmaps"],
        "postprocess_heatmap.py": ["--bg-heatmap-path", "--fg-heatmap-path", "--inv-heatmap-save-path",
                                   "--object-heatmap-path", "--save-dir", "--stack-heatmap-save-path"],
    }[name]
    import importlib
    mod = importlib.import_module("agenda_b200." + name[:-3])
    ours = {opt for action in mod.build_parser()._actions for opt in action.option_strings}
    assert set(recorded) <= ours
    ref_path = os.path.join("/root/reference/data_generation", name)
    if os.path.exists(ref_path):
        assert _flags(open(ref_path).read()) == recorded


def test_cli_defaults_match_reference():
    from agenda_b200 import data_generation, postprocess_heatmap
    a = data_generation.parse_args([])
    assert a.save_dir == "Data/Synthetic" and a.num_images == 10000 and a.image_size == 112
    assert a.prompt == "An aerial view image with {} cars in {} Utah"
    assert a.initialize_token == ["cars", "Utah", "New Zealand"] and a.word_token_heatmaps is None
    b = postprocess_heatmap.parse_args([])
    assert b.stack_heatmap_save_path == "daam_stack_heatmaps" and b.inv_heatmap_save_path == "daam_inv_heatmaps"


class _FakeTokenizer:
    def tokenize(self, s):
        return s.split()


def test_trace_shim_plumbing_cpu():
    """Install/restore of processors and word->row lookup, without running any kernel."""
    from agenda_b200.sd_attention import AttentionStack, BlockSpec
    from agenda_b200.trace import GlobalHeatMap, trace
    stack = AttentionStack([BlockSpec("down0", 4, 80, 2), BlockSpec("mid", 2, 80, 2)], context_dim=16)

    class Pipe:
        unet = stack
        tokenizer = _FakeTokenizer()

    sentinel = object()
    for m in list(stack.attn1) + list(stack.attn2):
        m.set_processor(sentinel)
    with trace(Pipe(), tokens=[2, 5], latent_hw=4, mode="daam") as trc:
        hooked = [m for m in list(stack.attn1) + list(stack.attn2) if m.processor is trc.hooker]
        assert len(hooked) == 1 and hooked[0] is stack.attn2[0]  # cross-attention of the non-mid block only
        with pytest.raises(RuntimeError, match="No heat maps found"):
            trc.compute_global_heat_map()
    assert all(m.processor is sentinel for m in list(stack.attn1) + list(stack.attn2))
    with trace(Pipe(), latent_hw=4) as trc:
        assert all(m.processor is trc.hooker for m in list(stack.attn1) + list(stack.attn2))
    g = GlobalHeatMap(torch.arange(3 * 4).reshape(3, 2, 2).float(), {1: 0, 2: 1, 3: 2}, _FakeTokenizer(), "an aerial cars view")
    assert g.token_indices("cars") == [3] and g.token_indices("aerial cars") == [2, 3]
    assert torch.equal(g.compute_word_heat_map("aerial cars").heatmap, (g.heat_maps[1] + g.heat_maps[2]) / 2)
    with pytest.raises(ValueError):
        g.token_indices("truck")


def test_processor_autograd_routing_is_decided_on_every_input():
    """Training mode (finetune_sd_token.py:754-757): with a frozen UNet only the prompt embedding requires grad.  The
    processor must route a call through its autograd path when ANY of hidden_states / encoder_hidden_states / to_q,
    to_k, to_v parameters requires grad (host logic only; the GPU tests check the gradients)."""
    from agenda_b200 import UNetCrossAttentionHooker
    from agenda_b200.sd_attention import SDAttention
    attn = SDAttention(64, 32, 2, 32)
    x, c = torch.zeros(1, 4, 64), torch.zeros(1, 3, 32)
    wants = UNetCrossAttentionHooker._wants_grad
    assert wants(attn, x, c)                                    # fresh nn.Linear weights require grad
    for p in attn.parameters():
        p.requires_grad_(False)
    assert not wants(attn, x, c) and not wants(attn, x, None)
    assert wants(attn, x, c.clone().requires_grad_(True))       # learned-token embedding
    assert wants(attn, x.clone().requires_grad_(True), None)    # activations downstream of a cross-attention layer
    attn.to_out[0].weight.requires_grad_(True)                  # to_out runs outside the attention core
    assert not wants(attn, x, c)
    # a CPU tensor never reaches a kernel silently, with or without a graph
    proc = UNetCrossAttentionHooker(is_train=True)
    with pytest.raises(RuntimeError):
        proc(attn, x.clone().requires_grad_(True))


@pytest.mark.parametrize("tokens,b_first", [([2, 5], 0), (None, 1), ([7, 7, 0], 1)])
def test_cross_attention_backward_formulas_cpu(tokens, b_first):
    """The gradient formulas of the training-mode cross-attention (agenda_b200/autograd.py docstring; the CUDA kernel
    agenda_attn_cross_bwd implements the same ones) against autograd of the oracle's attention + head-mean maps
    (hook.py:104-115, 28-56), on CPU in fp32: dq, dk, dv with gradients arriving through the output AND the maps."""
    from types import SimpleNamespace
    from agenda_b200.autograd import CrossAttentionHeatFn
    from oracle import hook_oracle as O
    g = torch.Generator().manual_seed(11)
    B, N, M, H, d = 2, 16, 9, 2, 8
    q, k, v = (torch.randn(B, s, H * d, generator=g) for s in (N, M, M))
    go = torch.randn(B, N, H * d, generator=g)
    T = M if tokens is None else len(tokens)
    gm = torch.randn(B - b_first, T, N, generator=g)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    out, p = O.attention_core(qr, kr, vr, H)
    maps = p.reshape(B, H, N, M)[b_first:].mean(1).permute(0, 2, 1)[:, list(range(M)) if tokens is None else tokens]
    torch.autograd.backward((out, maps), (go, gm))
    ctx = SimpleNamespace(saved_tensors=(q, k, v), heads=H, scale=d ** -0.5, b_first=b_first, token_idx=tokens)
    dq, dk, dv = CrossAttentionHeatFn._backward_torch(ctx, go, gm)[:3]
    for a, b_ in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert (a - b_).abs().max().item() < 1e-5 * max(1.0, b_.abs().max().item())


def test_fused_qkv_weight_declines_on_cpu_modules():
    """The one-GEMM q/k/v projection (and the scale fold into W_q) only applies to CUDA bf16 Linear layers; everything
    else must decline with a (None, False) pair so that the processor falls through to the separate projections."""
    from agenda_b200 import UNetCrossAttentionHooker
    from agenda_b200.sd_attention import SDAttention
    proc = UNetCrossAttentionHooker(is_train=False)
    assert proc._fused_qkv_weight(SDAttention(320, None, 8, 40)) == (None, False)            # fp32 on CPU
    assert proc._fused_qkv_weight(SDAttention(320, None, 8, 40).bfloat16()) == (None, False)  # bf16 but not CUDA
