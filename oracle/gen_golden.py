"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by EXECUTING THE REFERENCE.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py

* hook_*.npz : `/root/reference/data_generation/hook.py` is imported unmodified (with oracle/ref_stub providing
  the `diffusers` names it imports) and `UNetCrossAttentionHooker` is driven through stub `Attention` modules;
  inputs, weights, per-call outputs, recorded maps and `compute_global_heat_map()` are stored.
* post_*.npz : the numpy/PIL expressions of `data_generation.py:82-85` and `postprocess_heatmap.py:44-46`
  are executed literally on seeded inputs.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data_generation"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference_hook():
    sys.path.insert(0, os.path.join(HERE, "ref_stub"))
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import hook  # noqa: the reference file itself
    from diffusers.models.attention_processor import Attention
    return hook, Attention


def gen_hook(hook, Attention):
    torch.manual_seed(1234)
    ctx_dim, M = 32, 77
    # (name, h=w, heads, dim_head): every SD-1.x / SD-2.x head dim, small N so fixtures stay small
    layers = [("d40", 16, 2, 40), ("d80", 8, 2, 80), ("d160", 4, 1, 160), ("d64", 8, 2, 64), ("d40h4", 8, 4, 40)]
    for is_train in (False, True):
        hooker = hook.UNetCrossAttentionHooker(is_train=is_train, latent_hw=16)
        blob = {}
        B = 2
        ctx = torch.randn(B, M, ctx_dim)
        blob["ctx"] = ctx.numpy()
        for name, hw, heads, dh in (layers if not is_train else layers[:1]):
            C = heads * dh
            cross = Attention(C, cross_attention_dim=ctx_dim, heads=heads, dim_head=dh)
            selfa = Attention(C, heads=heads, dim_head=dh)
            with torch.no_grad():
                for mod in (cross, selfa):  # larger logits than default init => non-trivial softmax
                    mod.to_q.weight.mul_(4.0)
                    mod.to_k.weight.mul_(4.0)
            hs = torch.randn(B, hw * hw, C)
            with torch.no_grad():
                o_self = hooker(selfa, hs)
                n_before = len(hooker.cross_attn_maps)
                o_cross = hooker(cross, hs, encoder_hidden_states=ctx)
            assert len(hooker.cross_attn_maps) == n_before + 1
            for tag, mod, o in (("self", selfa, o_self), ("cross", cross, o_cross)):
                p = f"{name}_{tag}_"
                blob[p + "wq"] = mod.to_q.weight.detach().numpy()
                blob[p + "wk"] = mod.to_k.weight.detach().numpy()
                blob[p + "wv"] = mod.to_v.weight.detach().numpy()
                blob[p + "wo"] = mod.to_out[0].weight.detach().numpy()
                blob[p + "bo"] = mod.to_out[0].bias.detach().numpy()
                blob[p + "out"] = o.numpy()
            blob[f"{name}_hs"] = hs.numpy()
            blob[f"{name}_heads"] = np.int32(heads)
            blob[f"{name}_maps"] = hooker.cross_attn_maps[-1].numpy()
        blob["global"] = hooker.compute_global_heat_map().numpy()
        blob["layer_names"] = np.array([l[0] for l in (layers if not is_train else layers[:1])])
        np.savez_compressed(os.path.join(OUT, f"hook_call_{'train' if is_train else 'infer'}.npz"), **blob)

    # aggregation alone at the real latent size: scale x1, x2, x4, x8 (hook.py:59-81)
    torch.manual_seed(99)
    hooker = hook.UNetCrossAttentionHooker(is_train=False, latent_hw=64)
    blob = {}
    for i, hw in enumerate((64, 32, 16, 8, 32, 16)):
        m = torch.rand(1, 5, hw, hw) ** 4  # peaked, non-negative like attention probabilities
        hooker.cross_attn_maps.append(m)
        blob[f"map{i}"] = m.numpy()
    blob["global"] = hooker.compute_global_heat_map().numpy()
    hooker.clear()
    try:
        hooker.compute_global_heat_map()
        blob["empty_error"] = np.array("")
    except RuntimeError as e:
        blob["empty_error"] = np.array(str(e))
    # SD-2.1-768 latent (96) from 48/24/12
    hooker = hook.UNetCrossAttentionHooker(is_train=False, latent_hw=96)
    for i, hw in enumerate((96, 48, 24, 12)):
        m = torch.rand(1, 2, hw, hw) ** 4
        hooker.cross_attn_maps.append(m)
        blob[f"l96_map{i}"] = m.numpy()
    blob["l96_global"] = hooker.compute_global_heat_map().numpy()
    np.savez_compressed(os.path.join(OUT, "hook_global.npz"), **blob)


def gen_post():
    from PIL import Image
    rng = np.random.default_rng(7)
    blob = {}
    heats = []
    for i in range(6):
        h = rng.random((64, 64), dtype=np.float32) ** 3
        if i == 1:
            h *= 1e-6  # tiny dynamic range: the +1e-8 matters
        if i == 2:
            h[:] = 0.25  # constant map: 0/1e-8
        if i == 3:
            h = h * 40 - 7  # negative values
        heats.append(h)
    heats = np.stack(heats)
    blob["heat"] = heats
    u8, png = [], []
    for h in heats:
        object_daam_heatmap = h
        # data_generation.py:82
        object_daam_heatmap = (object_daam_heatmap - object_daam_heatmap.min()) / (object_daam_heatmap.max() - object_daam_heatmap.min() + 1e-8) * 255
        # data_generation.py:84-85
        word_DAAM_heat_pil = Image.fromarray(object_daam_heatmap.astype(np.uint8))
        u8.append(object_daam_heatmap.astype(np.uint8))
        word_DAAM_heat_pil = word_DAAM_heat_pil.resize((112, 112))
        png.append(np.asarray(word_DAAM_heat_pil))
    blob["u8"] = np.stack(u8)
    blob["png112"] = np.stack(png)
    # other target sizes exercised by --image-size
    blob["png64to200"] = np.asarray(Image.fromarray(u8[0]).resize((200, 200)))
    blob["png64to48"] = np.asarray(Image.fromarray(u8[0]).resize((48, 48)))
    # postprocess_heatmap.py:44-46
    obj_heatmap_arr, fg_heatmap_arr, bg_heatmap_arr = png[0], png[4], png[5]
    inv_bg_heatmap_arr = 255 - bg_heatmap_arr
    stack_heatmap_arr = np.stack([obj_heatmap_arr, fg_heatmap_arr, inv_bg_heatmap_arr], axis=-1)
    assert Image.fromarray(stack_heatmap_arr).mode == "RGB"
    blob["stack"] = stack_heatmap_arr
    blob["inv_bg"] = inv_bg_heatmap_arr
    np.savez_compressed(os.path.join(OUT, "post.npz"), **blob)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    hook, Attention = _import_reference_hook()
    gen_hook(hook, Attention)
    gen_post()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
