"""Host-buffer API of the post-processing kernels (numpy in, numpy out; H2D -> CUDA kernel -> D2H), plus the
reference's fixed-size box rule.  This is the layer the CLI mirrors and INTEGRATION.md's reference-side stubs call.

There is no CPU implementation here: every function needs a CUDA device and raises otherwise.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ops


def _cuda(a: np.ndarray, dtype) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("agenda_b200.postprocess needs a CUDA device (no CPU fallback exists)")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda(non_blocking=True)


def heatmap_to_png_array(heat: np.ndarray, image_size: int = 112) -> np.ndarray:
    """data_generation.py:82-85 for one map or a batch: fp32 [...,h,w] -> u8 [...,image_size,image_size]."""
    return ops.heat_to_u8_image(_cuda(heat, np.float32), image_size).cpu().numpy()


def stack_heatmaps(obj: np.ndarray, fg: np.ndarray, bg: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """postprocess_heatmap.py:44-46: (stack [...,H,W,3], inverted bg [...,H,W]), u8."""
    st, inv = ops.stack_heatmaps_u8(_cuda(obj, np.uint8), _cuda(fg, np.uint8), _cuda(bg, np.uint8))
    return st.cpu().numpy(), inv.cpu().numpy()


def ccl_bbox(heat: np.ndarray, thr: float = 0.5, max_boxes: int = 256, want_labels: bool = True):
    """fp32 [n,H,W] (or [H,W]) -> (labels int32 or None, list of int32 [K_i,5] boxes x,y,w,h,area)."""
    h = np.asarray(heat, dtype=np.float32)
    single = h.ndim == 2
    labels, counts, boxes = ops.ccl_bbox(_cuda(h[None] if single else h, np.float32), thr, max_boxes, want_labels)
    counts = counts.cpu().numpy()
    boxes = boxes.cpu().numpy()
    out_boxes = [boxes[i, :min(int(c), max_boxes)].copy() for i, c in enumerate(counts)]
    lab = labels.cpu().numpy() if want_labels else None
    if single:
        return (lab[0] if want_labels else None), out_boxes[0]
    return lab, out_boxes


def fixed_size_boxes(boxes_xywh: np.ndarray, box_px: float = 42.36, image_size: Tuple[int, int] = (112, 112)) -> np.ndarray:
    """The reference's box geometry rule (data_annotation/refine_label.py:17-18,58-113): every detection becomes a
    `box_px` square around its centre; boxes whose centre is within margin = box_px/2 - 1 of an edge are first
    completed to full size away from that edge (they were clipped by the image border), then the square is
    clipped to the image.  In: [K,4+] x,y,w,h (COCO top-left, Data/README.md:7).  Out: float64 [K,4] x,y,w,h.
    Host arithmetic in Python float64, exactly as the reference."""
    b = np.asarray(boxes_xywh, dtype=np.float64)
    out = np.zeros((b.shape[0], 4), dtype=np.float64)
    margin = box_px / 2 - 1
    W, H = image_size
    for i in range(b.shape[0]):
        l, t = b[i, 0], b[i, 1]
        r, bt = l + b[i, 2], t + b[i, 3]
        xc, yc = (l + r) / 2, (t + bt) / 2
        if xc < margin:
            l_full, r_full = r - box_px, r
        elif xc > W - margin:
            l_full, r_full = l, l + box_px
        else:
            l_full, r_full = l, r
        if yc < margin:
            t_full, b_full = bt - box_px, bt
        elif yc > H - margin:
            t_full, b_full = t, t + box_px
        else:
            t_full, b_full = t, bt
        xcf, ycf = (l_full + r_full) / 2, (t_full + b_full) / 2
        l2 = max(0, xcf - box_px / 2)
        t2 = max(0, ycf - box_px / 2)
        r2 = min(xcf + box_px / 2, W - 1)
        b2 = min(ycf + box_px / 2, H - 1)
        out[i] = (l2, t2, r2 - l2, b2 - t2)
    return out
