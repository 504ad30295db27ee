"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy / torch-CPU fp32) of the AGenDA heat-map hot path.

This file is the *checker*: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  Nothing under `agenda_b200/` does (the product path fails loudly when
the CUDA library is missing, it never falls back to this).

Every function cites the reference lines it restates (paths are into /root/reference/).  Pinning status:

  * attention processor, `_unravel_attn`, `compute_global_heat_map`  — PINNED: `tests/golden/hook_*.npz` were
    produced by executing the reference's `data_generation/hook.py` unmodified (oracle/gen_golden.py) and
    `tests/test_oracle.py` checks this restatement against them.
  * normalise/quantise/resize (`data_generation.py:82-85`) and invert/stack (`postprocess_heatmap.py:44-46`) —
    PINNED: those reference lines are plain numpy/PIL calls; the golden vectors were produced by running those
    exact expressions (numpy + PIL are in the image), and `pil_resize_bicubic_u8` (our restatement of PIL's
    fixed-point resampler, needed because the CUDA kernel must reproduce it) is checked against PIL itself.
  * threshold / connected components / bbox — PARITY UNPINNED: the reference contains no such code
    (SURVEY.md §0 D3).  The spec is restated here over `scipy.ndimage.label`.
  * DAAM-mode aggregation (`daam` is an un-vendored, un-pinned PyPI dependency, requirements.txt:4) —
    PARITY UNPINNED: restated from daam's published algorithm.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------------------
# a2: UNetCrossAttentionHooker.__call__  (data_generation/hook.py:83-122) + the diffusers-0.21.2 helpers it calls
# --------------------------------------------------------------------------------------------------------------


def head_to_batch_dim(x: torch.Tensor, heads: int) -> torch.Tensor:
    """diffusers Attention.head_to_batch_dim as used at hook.py:104-106: [B,S,H*d] -> [B*H,S,d], row = b*H+h."""
    b, s, c = x.shape
    return x.reshape(b, s, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, s, c // heads)


def batch_to_head_dim(x: torch.Tensor, heads: int) -> torch.Tensor:
    """diffusers Attention.batch_to_head_dim as used at hook.py:115."""
    bh, s, d = x.shape
    return x.reshape(bh // heads, heads, s, d).permute(0, 2, 1, 3).reshape(bh // heads, s, d * heads)


def attention_probs(q: torch.Tensor, k: torch.Tensor, scale: float, mask: torch.Tensor = None) -> torch.Tensor:
    """attn.get_attention_scores (hook.py:108): softmax(scale * q k^T [+ attention_mask]) over the key axis, fp32.
    mask: what attn.prepare_attention_mask returned (hook.py:92), additive, [B*H, 1 | N, M] (baddbmm input, beta = 1)."""
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if mask is not None:
        s = s + mask.float()
    return s.softmax(dim=-1)


def unravel_attn(probs: torch.Tensor, heads: int, is_train: bool) -> torch.Tensor:
    """hook.py:28-56.  probs [B*H, N, M] -> [B', M, h, w]: drop the unconditional half of the batch-head axis
    when not training (hook.py:48-49), regroup per (batch, head) and MEAN over heads (hook.py:54-55)."""
    bh, n, m = probs.shape
    h = w = int(math.sqrt(n))
    x = probs.permute(2, 0, 1).reshape(m, bh, h, w)
    if not is_train:
        x = x[:, bh // 2:]
    x = x.permute(1, 0, 2, 3)  # [B'*H, M, h, w]
    x = x.reshape(x.shape[0] // heads, heads, m, h, w)
    return x.mean(dim=1)


def processor_call(hidden_states, encoder_hidden_states, wq, wk, wv, wo, bo, heads: int, is_train: bool, mask=None):
    """hook.py:83-122 with explicit weights (to_q/to_k/to_v bias-free, to_out[0] with bias, dropout p=0).

    Returns (out [B,N,C], maps [B',M,h,w] or None for self-attention)."""
    hs = hidden_states.float()
    is_cross = encoder_hidden_states is not None
    ehs = encoder_hidden_states.float() if is_cross else hs
    q = hs @ wq.float().t()
    k = ehs @ wk.float().t()
    v = ehs @ wv.float().t()
    d = q.shape[-1] // heads
    qh, kh, vh = (head_to_batch_dim(t, heads) for t in (q, k, v))
    p = attention_probs(qh, kh, d ** -0.5, mask)
    maps = unravel_attn(p, heads, is_train) if is_cross else None
    o = batch_to_head_dim(torch.bmm(p, vh), heads)
    out = o @ wo.float().t() + bo.float()
    return out, maps


def attention_core(q, k, v, heads: int):
    """The part of hook.py:104-115 the fused CUDA kernels replace: [B,N,C],[B,M,C],[B,M,C] -> O [B,N,C], P."""
    d = q.shape[-1] // heads
    qh, kh, vh = (head_to_batch_dim(t.float(), heads) for t in (q, k, v))
    p = attention_probs(qh, kh, d ** -0.5)
    return batch_to_head_dim(torch.bmm(p, vh), heads), p


# --------------------------------------------------------------------------------------------------------------
# a4: compute_global_heat_map  (data_generation/hook.py:59-81)
# --------------------------------------------------------------------------------------------------------------

_A = np.float32(-0.75)  # torch's bicubic constant (F.interpolate mode='bicubic', hook.py:72)


def _cubic_coeffs(t: np.ndarray) -> np.ndarray:
    """torch get_cubic_upsample_coefficients in fp32: taps for offsets -1, 0, +1, +2."""
    t = t.astype(np.float32)
    one = np.float32(1)

    def c1(x):
        return ((_A + np.float32(2)) * x - (_A + np.float32(3))) * x * x + one

    def c2(x):
        return ((_A * x - np.float32(5) * _A) * x + np.float32(8) * _A) * x - np.float32(4) * _A

    return np.stack([c2(t + one), c1(t), c1(one - t), c2((one - t) + one)], axis=-1).astype(np.float32)


def _bicubic_axis_tables(n_in: int, n_out: int):
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = scale * (dst + np.float32(0.5)) - np.float32(0.5)  # align_corners=False, cubic: no clamp at 0
    i0 = np.floor(src)
    t = (src - i0).astype(np.float32)
    idx = i0.astype(np.int64)[:, None] + np.arange(-1, 3)[None, :]
    idx = np.clip(idx, 0, n_in - 1)  # border replicate
    return idx, _cubic_coeffs(t)


def bicubic_upsample(x: np.ndarray, size: int) -> np.ndarray:
    """F.interpolate(x, size=(size,size), mode='bicubic') (hook.py:72; align_corners=False), separable fp32."""
    x = np.asarray(x, dtype=np.float32)
    h, w = x.shape[-2:]
    if h == size and w == size:
        return x.copy()  # torch returns the input values bit-exactly at scale 1
    iy, wy = _bicubic_axis_tables(h, size)
    ix, wx = _bicubic_axis_tables(w, size)
    # horizontal taps within each source row, then combine the 4 source rows (torch's CPU kernel order)
    rows = np.zeros(x.shape[:-1] + (size,), dtype=np.float32)
    for j in range(4):
        rows += x[..., ix[:, j]] * wx[:, j]
    out = np.zeros(x.shape[:-2] + (size, size), dtype=np.float32)
    for j in range(4):
        out += rows[..., iy[:, j], :] * wy[:, j][:, None]
    return out


def global_heat_map(maps, latent_hw: int = 64) -> np.ndarray:
    """hook.py:59-81: every stored [B',M,h,w] map -> bicubic to latent_hw² -> clamp(min=0) -> mean over the
    list.  Raises RuntimeError('No heat maps found.') on an empty list (hook.py:74-77)."""
    if len(maps) == 0:
        raise RuntimeError('No heat maps found.')
    acc = None
    for m in maps:
        m = m.detach().cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
        up = np.maximum(bicubic_upsample(m, latent_hw), np.float32(0))
        acc = up.astype(np.float32) if acc is None else acc + up
    return (acc / np.float32(len(maps))).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------
# a5/a6: DAAM-mode aggregation (daam is NOT in the reference tree: PARITY UNPINNED, restated from memory of
# castorini/daam's trace.py / heatmap.py; call sites data_generation.py:57,64,74-77)
# --------------------------------------------------------------------------------------------------------------


def daam_global_heat_map(per_layer_head_sums, latent_hw: int = 64, n_rows: int | None = None) -> np.ndarray:
    """DAAM keeps, per hooked layer, the sum over timesteps of the per-head conditional maps [H, M, h, w]
    (layers with h*8 == latent are skipped by DAAM's `factor==8` rule — the caller filters).  At the end every
    (layer, head) map is bicubic-upsampled + clamped, and the mean over all of them is taken."""
    ups = []
    for m in per_layer_head_sums:
        m = np.asarray(m, dtype=np.float32)
        for hd in range(m.shape[0]):
            ups.append(np.maximum(bicubic_upsample(m[hd], latent_hw), np.float32(0)))
    if not ups:
        raise RuntimeError('No heat maps found. Did you forget to call `with trace(...)` during generation?')
    out = np.mean(np.stack(ups, 0), axis=0, dtype=np.float32)
    return out[:n_rows] if n_rows is not None else out


def word_heat_map(global_map: np.ndarray, token_indices) -> np.ndarray:
    """daam GlobalHeatMap.compute_word_heat_map (data_generation.py:74-77): mean over the word's token rows."""
    return np.asarray(global_map, dtype=np.float32)[list(token_indices)].mean(axis=0, dtype=np.float32)


# --------------------------------------------------------------------------------------------------------------
# a7: normalise + quantise + resize  (data_generation/data_generation.py:82-85)
# --------------------------------------------------------------------------------------------------------------


def normalize_u8(heat: np.ndarray) -> np.ndarray:
    """data_generation.py:82 + the astype at :84 — verbatim numpy on an fp32 map."""
    h = np.asarray(heat, dtype=np.float32)
    h = (h - h.min()) / (h.max() - h.min() + 1e-8) * 255
    return h.astype(np.uint8)


def pil_resize(img_u8: np.ndarray, size: int) -> np.ndarray:
    """data_generation.py:84-85 verbatim: Image.fromarray(u8).resize((size,size)) (PIL default = BICUBIC)."""
    from PIL import Image
    return np.asarray(Image.fromarray(np.asarray(img_u8, dtype=np.uint8)).resize((size, size)))


_PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """Restatement of Pillow's precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) for the
    BICUBIC filter over the whole axis: returns (bounds [out,2] = (xmin, count), kk int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_pil_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << _PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PIL_PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pil_clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> _PIL_PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Restatement of PIL's two-pass 8-bit resampler (horizontal pass into a u8 intermediate, then vertical),
    which is what data_generation.py:85 executes for mode 'L'.  Checked against PIL in tests/test_oracle.py."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape
    cur = img
    if out_w != w:
        b, kk = pil_bicubic_coeffs(w, out_w)
        tmp = np.zeros((h, out_w), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = b[xx]
            acc = (cur[:, x0:x0 + n].astype(np.int64) * kk[xx, :n].astype(np.int64)).sum(axis=1)
            tmp[:, xx] = _pil_clip8(acc + (1 << (_PIL_PRECISION_BITS - 1)))
        cur = tmp
    if out_h != h:
        b, kk = pil_bicubic_coeffs(h, out_h)
        tmp = np.zeros((out_h, cur.shape[1]), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = b[yy]
            acc = (cur[y0:y0 + n, :].astype(np.int64) * kk[yy, :n, None].astype(np.int64)).sum(axis=0)
            tmp[yy] = _pil_clip8(acc + (1 << (_PIL_PRECISION_BITS - 1)))
        cur = tmp
    return cur


def heat_to_png_array(heat: np.ndarray, image_size: int = 112) -> np.ndarray:
    """data_generation.py:78-85 end to end for one word: fp32 [L,L] -> u8 [image_size,image_size]."""
    return pil_resize(normalize_u8(heat), image_size)


# --------------------------------------------------------------------------------------------------------------
# a8: invert + stack  (data_generation/postprocess_heatmap.py:44-46)
# --------------------------------------------------------------------------------------------------------------


def stack_heatmaps(obj: np.ndarray, fg: np.ndarray, bg: np.ndarray):
    """postprocess_heatmap.py:44-46 verbatim: returns (stack [H,W,3] u8, inv_bg [H,W] u8)."""
    inv_bg = 255 - np.asarray(bg, dtype=np.uint8)
    return np.stack([np.asarray(obj, dtype=np.uint8), np.asarray(fg, dtype=np.uint8), inv_bg], axis=-1), inv_bg


# --------------------------------------------------------------------------------------------------------------
# a9: threshold / CCL / bbox — NOT IN THE REFERENCE (PARITY UNPINNED; spec = SURVEY.md §8 a9)
# --------------------------------------------------------------------------------------------------------------


def ccl_bbox(heat: np.ndarray, thr: float = 0.5):
    """n = (h-min)/((max-min)+1e-8f) in fp32; mask = n > thr; 4-connectivity; labels 1..K numbered in raster
    order of each component's first pixel (scipy.ndimage.label's rule); boxes int32 [K,5] = x, y, w, h, area
    (COCO top-left convention, Data/README.md:7)."""
    from scipy import ndimage
    h = np.asarray(heat, dtype=np.float32)
    n = (h - h.min()) / ((h.max() - h.min()) + np.float32(1e-8))
    mask = n > np.float32(thr)
    labels, k = ndimage.label(mask, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    labels = labels.astype(np.int32)
    boxes = np.zeros((k, 5), dtype=np.int32)
    for i, sl in enumerate(ndimage.find_objects(labels)):
        ys, xs = sl
        boxes[i] = (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start,
                    int((labels[sl] == i + 1).sum()))
    return labels, boxes


def synthetic_heatmaps(n: int, size: int = 512, seed: int = 0) -> np.ndarray:
    """BASELINE.json config 5 generator: sum of K in [0,40] Gaussian blobs (sigma in [4,16] px,
    amplitude U(0.2,1)) over a 0.02*U(0,1) noise floor, fp32 [n,size,size]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.empty((n, size, size), dtype=np.float32)
    for i in range(n):
        img = (0.02 * rng.random((size, size))).astype(np.float32)
        for _ in range(int(rng.integers(0, 41))):
            cx, cy = rng.uniform(0, size, 2)
            sig = rng.uniform(4, 16)
            amp = rng.uniform(0.2, 1.0)
            r = int(4 * sig) + 1
            x0, x1 = max(0, int(cx) - r), min(size, int(cx) + r + 1)
            y0, y1 = max(0, int(cy) - r), min(size, int(cy) + r + 1)
            img[y0:y1, x0:x1] += (amp * np.exp(-((xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2)
                                               / (2 * sig * sig))).astype(np.float32)
        out[i] = img
    return out


# --------------------------------------------------------------------------------------------------------------
# N2: the reference's fixed-size box rule  (data_annotation/refine_label.py:17-18,58-113), one box, Python floats
# --------------------------------------------------------------------------------------------------------------


def fixed_size_box_rule(l, t, r, b, bboxes_size_px=42.36, image_size=(112, 112)):
    """refine_label.py:58-113 restated for one detection (l, t, r, b); rgb_image.size == image_size.  Returns the COCO
    box (x, y, w, h) the reference appends at refine_label.py:126-135.  (The reference hard-codes 42.36 in the
    edge-completion branch, :78-100, and uses bboxes_size_px for the final square, :105-109.)"""
    margin_px = bboxes_size_px / 2 - 1
    x_c_bbox = (l + r) / 2
    y_c_bbox = (t + b) / 2
    v = 'left' if x_c_bbox < margin_px else ('right' if x_c_bbox > image_size[0] - margin_px else None)
    h = 'top' if y_c_bbox < margin_px else ('bottom' if y_c_bbox > image_size[1] - margin_px else None)
    if v == 'left':
        r_full, l_full = r, r - 42.36
    elif v == 'right':
        l_full, r_full = l, l + 42.36
    else:
        l_full, r_full = l, r
    if h == 'top':
        b_full, t_full = b, b - 42.36
    elif h == 'bottom':
        t_full, b_full = t, t + 42.36
    else:
        t_full, b_full = t, b
    xc, yc = (l_full + r_full) / 2, (t_full + b_full) / 2
    l2 = max(0, xc - bboxes_size_px / 2)
    t2 = max(0, yc - bboxes_size_px / 2)
    r2 = min(xc + bboxes_size_px / 2, image_size[0] - 1)
    b2 = min(yc + bboxes_size_px / 2, image_size[1] - 1)
    return l2, t2, r2 - l2, b2 - t2


def coco_boxes_from_heat(heat: np.ndarray, thr: float = 0.5, image_size: int = 112):
    """a9 boxes of one fp32 heat map [L, L] -> the annotation boxes the generation driver writes: CCL boxes scaled from
    the map grid to the image grid (x S/L), then the fixed-size rule.  List of (x, y, w, h) float tuples."""
    _, boxes = ccl_bbox(heat, thr)
    s = image_size / heat.shape[-1]
    out = []
    for x, y, w, h, _ in boxes.tolist():
        out.append(fixed_size_box_rule(x * s, y * s, x * s + w * s, y * s + h * s, image_size=(image_size, image_size)))
    return out
