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
    float64 throughout, the reference's own operations in the same order (it computes in Python floats), over all
    boxes at once."""
    b = np.asarray(boxes_xywh, dtype=np.float64)
    if b.size == 0:
        return np.zeros((0, 4), dtype=np.float64)
    if b.ndim == 1:
        b = b[None]
    margin = box_px / 2 - 1
    W, H = image_size
    l, t = b[:, 0], b[:, 1]
    r, bt = l + b[:, 2], t + b[:, 3]
    xc, yc = (l + r) / 2, (t + bt) / 2
    left, right = xc < margin, (xc >= margin) & (xc > W - margin)
    top, bottom = yc < margin, (yc >= margin) & (yc > H - margin)
    l_full = np.where(left, r - box_px, l)
    r_full = np.where(left, r, np.where(right, l + box_px, r))
    t_full = np.where(top, bt - box_px, t)
    b_full = np.where(top, bt, np.where(bottom, t + box_px, bt))
    xcf, ycf = (l_full + r_full) / 2, (t_full + b_full) / 2
    l2 = np.maximum(0, xcf - box_px / 2)
    t2 = np.maximum(0, ycf - box_px / 2)
    r2 = np.minimum(xcf + box_px / 2, W - 1)
    b2 = np.minimum(ycf + box_px / 2, H - 1)
    return np.stack([l2, t2, r2 - l2, b2 - t2], axis=1)


def boxes_to_image_coords(boxes_xywh: np.ndarray, map_size: Tuple[int, int], image_size: Tuple[int, int]) -> np.ndarray:
    """CCL boxes live on the heat map's pixel grid (L x L latent resolution); the data set's images and annotations are
    `image_size` (112 x 112, data_generation.py:21,60; Data/README.md:7).  Pixel p of the map covers [p, p+1) * S/L in
    the image: x, y, w, h scale by S/L (float64).  In [K,4+] -> out float64 [K,4]."""
    b = np.asarray(boxes_xywh, dtype=np.float64)
    if b.size == 0:
        return np.zeros((0, 4), dtype=np.float64)
    sx, sy = image_size[0] / map_size[0], image_size[1] / map_size[1]
    return np.stack([b[:, 0] * sx, b[:, 1] * sy, b[:, 2] * sx, b[:, 3] * sy], axis=1)


COCO_CATEGORIES = [{"id": 1, "name": "small"}]   # refine_label.py:20-22 ("small" = the small-vehicle class, Data/README.md:7)


def coco_annotations(file_names, boxes_per_image, image_size: Tuple[int, int] = (112, 112), image_ids=None) -> dict:
    """COCO dict in the layout the reference writes for its pseudo-labels (refine_label.py:28-48,126-135): categories
    [{'id': 1, 'name': 'small'}], images [{id, file_name, width, height}], annotations [{iscrowd, category_id, image_id,
    bbox [x, y, w, h] (top-left, Data/README.md:7), area, label}] (+ a running 'id', which pycocotools needs)."""
    ids = list(range(len(file_names))) if image_ids is None else [int(i) for i in image_ids]
    images, anns = [], []
    for img_id, name, boxes in zip(ids, file_names, boxes_per_image):
        images.append({"id": img_id, "file_name": name, "width": int(image_size[0]), "height": int(image_size[1])})
        for bx in np.asarray(boxes, dtype=np.float64).reshape(-1, 4):
            x, y, w, h = (float(v) for v in bx)
            anns.append({"id": len(anns), "iscrowd": 0, "category_id": 1, "image_id": img_id, "bbox": [x, y, w, h],
                         "area": w * h, "label": 1})
    return {"categories": [dict(c) for c in COCO_CATEGORIES], "images": images, "annotations": anns}


def write_coco_json(path: str, coco: dict) -> None:
    import json
    with open(path, "w") as f:
        json.dump(coco, f, indent=4)


def ccl_boxes_to_coco_boxes(counts, boxes, map_size: int, image_size: int = 112, box_px: float = 42.36,
                            max_boxes: int = None):
    """Device output of the CCL kernel (counts int32 [n], boxes int32 [n,max_boxes,5]) -> per image the reference-style
    annotation boxes: scaled to the image grid, then the fixed-size rule.  Returns a list of float64 [K_i,4]."""
    counts = np.asarray(counts.cpu() if hasattr(counts, "cpu") else counts)
    boxes = np.asarray(boxes.cpu() if hasattr(boxes, "cpu") else boxes)
    cap = boxes.shape[1] if max_boxes is None else max_boxes
    out = []
    for i, c in enumerate(counts):
        raw = boxes[i, :min(int(c), cap), :4]
        out.append(fixed_size_boxes(boxes_to_image_coords(raw, (map_size, map_size), (image_size, image_size)), box_px,
                                    (image_size, image_size)))
    return out
