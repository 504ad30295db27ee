"""CPU: pins oracle/hook_oracle.py against the golden vectors produced by executing the reference itself
(oracle/gen_golden.py: hook.py unmodified; numpy/PIL lines of data_generation.py / postprocess_heatmap.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import hook_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("mode", ["infer", "train"])
def test_processor_call_matches_reference_hook(golden_dir, mode):
    g = _load(golden_dir, f"hook_call_{mode}.npz")
    ctx = torch.from_numpy(g["ctx"])
    maps = []
    for name in g["layer_names"]:
        heads = int(g[f"{name}_heads"])
        hs = torch.from_numpy(g[f"{name}_hs"])
        for tag in ("self", "cross"):
            w = {k: torch.from_numpy(g[f"{name}_{tag}_{k}"]) for k in ("wq", "wk", "wv", "wo", "bo")}
            out, m = O.processor_call(hs, ctx if tag == "cross" else None, w["wq"], w["wk"], w["wv"], w["wo"],
                                      w["bo"], heads, is_train=(mode == "train"))
            np.testing.assert_allclose(out.numpy(), g[f"{name}_{tag}_out"], rtol=0, atol=2e-5)
            if tag == "cross":
                ref = g[f"{name}_maps"]
                assert m.shape == ref.shape
                assert ref.shape[0] == (2 if mode == "train" else 1)  # hook.py:48-49 drops the uncond half
                np.testing.assert_allclose(m.numpy(), ref, rtol=0, atol=2e-6)
                maps.append(m)
            else:
                assert m is None
    np.testing.assert_allclose(O.global_heat_map(maps, 16), g["global"], rtol=0, atol=2e-6)


def test_global_heat_map_matches_reference(golden_dir):
    g = _load(golden_dir, "hook_global.npz")
    maps = [g[f"map{i}"] for i in range(6)]
    np.testing.assert_allclose(O.global_heat_map(maps, 64), g["global"], rtol=0, atol=2e-6)
    maps96 = [g[f"l96_map{i}"] for i in range(4)]
    np.testing.assert_allclose(O.global_heat_map(maps96, 96), g["l96_global"], rtol=0, atol=2e-6)
    assert str(g["empty_error"]) == "No heat maps found."
    with pytest.raises(RuntimeError, match="No heat maps found."):
        O.global_heat_map([], 64)


def test_bicubic_known_taps():
    """SURVEY.md §8 a4 tap table (x2, phase t=0.75) and identity at scale 1."""
    c = O._cubic_coeffs(np.array([0.75], dtype=np.float32))[0]
    np.testing.assert_allclose(c, [-0.035156, 0.261719, 0.878906, -0.105469], atol=1e-6)
    x = np.random.default_rng(0).random((3, 8, 8), dtype=np.float32)
    assert np.array_equal(O.bicubic_upsample(x, 8), x)
    up = O.bicubic_upsample(x, 64)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x)[None], size=(64, 64), mode="bicubic")[0].numpy()
    np.testing.assert_allclose(up, ref, atol=2e-6)
    assert ref.min() < 0  # undershoot is real => the clamp at hook.py:72 is not a no-op


def test_normalize_resize_stack_match_reference_lines(golden_dir):
    g = _load(golden_dir, "post.npz")
    for i, h in enumerate(g["heat"]):
        u8 = O.normalize_u8(h)
        assert np.array_equal(u8, g["u8"][i])
        assert np.array_equal(O.pil_resize(u8, 112), g["png112"][i])
        assert np.array_equal(O.pil_resize_bicubic_u8(u8, 112, 112), g["png112"][i])  # our PIL restatement
    assert np.array_equal(O.pil_resize_bicubic_u8(g["u8"][0], 200, 200), g["png64to200"])
    assert np.array_equal(O.pil_resize_bicubic_u8(g["u8"][0], 48, 48), g["png64to48"])
    stack, inv = O.stack_heatmaps(g["png112"][0], g["png112"][4], g["png112"][5])
    assert np.array_equal(stack, g["stack"]) and np.array_equal(inv, g["inv_bg"])
    assert stack.shape == (112, 112, 3) and stack.dtype == np.uint8


def test_normalize_u8_truncates():
    h = np.array([[0.0, 1.0], [254.999 / 255.0, 0.5]], dtype=np.float32)
    # numpy>=2 (NEP 50): float32 + python 1e-8 stays float32, so (1.0 + 1e-8) == 1.0f and the max maps to 255
    assert O.normalize_u8(h).tolist() == [[0, 255], [254, 127]]


def test_pil_restatement_random_sizes():
    from PIL import Image
    rng = np.random.default_rng(3)
    for (h, w, oh, ow) in [(64, 64, 112, 112), (96, 96, 112, 112), (64, 64, 64, 64), (64, 64, 33, 97), (17, 40, 112, 5)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh)))
        assert np.array_equal(O.pil_resize_bicubic_u8(img, oh, ow), ref), (h, w, oh, ow)


def test_ccl_oracle_spec():
    h = np.zeros((6, 8), dtype=np.float32)
    h[0, 5:7] = 1.0          # component first met in raster order -> label 1
    h[1:3, 0:2] = 0.9        # label 2
    h[2, 2] = 0.8            # 4-connected to label 2 via (2,1)
    h[3, 3] = 0.7            # diagonal only -> its own component (4-connectivity)
    h[5, 7] = 0.51
    h[4, 4] = 0.5            # == thr after normalisation? n = 0.5/(1+1e-8) < 0.5 -> background (strict >)
    labels, boxes = O.ccl_bbox(h, 0.5)
    assert labels.max() == 4
    assert boxes.tolist() == [[5, 0, 2, 1, 2], [0, 1, 3, 2, 5], [3, 3, 1, 1, 1], [7, 5, 1, 1, 1]]
    assert labels[4, 4] == 0


def test_ccl_oracle_agrees_with_opencv():
    """Second, independent pin for the a9 spec (absent from the reference): OpenCV's
    connectedComponentsWithStats at 4-connectivity numbers components in the same raster order as scipy.ndimage.label
     and reports the same (x, y, w, h, area) — on config-5 style maps, salt-and-pepper noise and edge cases."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    maps = [m for m in O.synthetic_heatmaps(6, 256, seed=9)] + [rng.random((96, 160), dtype=np.float32) for _ in range(3)]
    maps += [np.zeros((40, 64), np.float32), np.ones((40, 64), np.float32), np.eye(48, dtype=np.float32)]
    for h in maps:
        for thr in (0.5, 0.2, 0.8):
            labels, boxes = O.ccl_bbox(h, thr)
            mn, mx = h.min(), h.max()
            mask = ((h - mn) / ((mx - mn) + np.float32(1e-8)) > np.float32(thr)).astype(np.uint8)
            n, lab, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=4, ltype=cv2.CV_32S)
            assert n - 1 == len(boxes)
            assert np.array_equal(lab, labels)
            assert np.array_equal(stats[1:].astype(np.int64), boxes.astype(np.int64))   # [x, y, w, h, area]


def test_synthetic_heatmaps_deterministic():
    from agenda_b200.synthetic import synthetic_heatmaps
    a = O.synthetic_heatmaps(2, 64, seed=0)
    b = O.synthetic_heatmaps(2, 64, seed=0)
    assert a.dtype == np.float32 and np.array_equal(a, b)
    assert np.array_equal(a, synthetic_heatmaps(2, 64, seed=0))  # product-side workload generator == oracle's
