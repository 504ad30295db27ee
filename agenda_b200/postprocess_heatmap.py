"""CLI mirror of the reference's `data_generation/postprocess_heatmap.py` (same flags, same outputs), with the
invert + stack arithmetic (postprocess_heatmap.py:44-46) running in `agenda_stack_heatmaps_u8` on the GPU.

    python -m agenda_b200.postprocess_heatmap --save-dir Data/Synthetic --object-heatmap-path daam_cars_heatmaps \
        --fg-heatmap-path daam_<fg>_heatmaps --bg-heatmap-path daam_<bg>_heatmaps

Difference from the reference, on purpose: files are paired by SORTED name (the reference zips three unsorted
`os.listdir` results, postprocess_heatmap.py:32-36, which only works when the three directories enumerate in the
same order), and a name missing from any of the three directories is an error instead of a silent mis-pairing.
"""
from __future__ import annotations

import argparse
import os

import numpy as np


# The reference's command line (postprocess_heatmap.py:8-17): flag -> default.  Names and defaults are the drop-in
# surface; all six take a string.
_REFERENCE_FLAGS = (
    ("--save-dir", "Data/Synthetic", "root that holds the heat-map directories"),
    ("--object-heatmap-path", None, "sub-directory with the object word's heat maps"),
    ("--fg-heatmap-path", None, "sub-directory with the foreground token's heat maps"),
    ("--bg-heatmap-path", None, "sub-directory with the background token's heat maps"),
    ("--stack-heatmap-save-path", "daam_stack_heatmaps", "sub-directory for the RGB stacks"),
    ("--inv-heatmap-save-path", "daam_inv_heatmaps", "sub-directory for the inverted background maps"),
)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Invert the background heat map and stack object / fg / bg into RGB.")
    for flag, default, text in _REFERENCE_FLAGS:
        parser.add_argument(flag, type=str, default=default, help=text)
    parser.add_argument("--batch", type=int, default=1024, help="images per GPU launch (not in the reference)")
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def main(argv=None):
    from PIL import Image
    from . import postprocess
    args = parse_args(argv)
    for name in ("object_heatmap_path", "fg_heatmap_path", "bg_heatmap_path"):
        if getattr(args, name) is None:
            raise SystemExit(f"--{name.replace('_', '-')} is required")
    obj_dir = os.path.join(args.save_dir, args.object_heatmap_path)
    fg_dir = os.path.join(args.save_dir, args.fg_heatmap_path)
    bg_dir = os.path.join(args.save_dir, args.bg_heatmap_path)
    stack_dir = os.path.join(args.save_dir, args.stack_heatmap_save_path)
    inv_dir = os.path.join(args.save_dir, args.inv_heatmap_save_path)
    os.makedirs(stack_dir, exist_ok=True)
    os.makedirs(inv_dir, exist_ok=True)

    names = sorted(os.listdir(obj_dir))
    for d in (fg_dir, bg_dir):
        missing = set(names) ^ set(os.listdir(d))
        if missing:
            raise SystemExit(f"heat-map directories do not hold the same files (e.g. {sorted(missing)[:3]})")

    def load(d, chunk):
        return np.stack([np.asarray(Image.open(os.path.join(d, n))) for n in chunk])

    for i in range(0, len(names), args.batch):
        chunk = names[i:i + args.batch]
        stack, inv = postprocess.stack_heatmaps(load(obj_dir, chunk), load(fg_dir, chunk), load(bg_dir, chunk))
        for k, n in enumerate(chunk):
            Image.fromarray(stack[k]).save(os.path.join(stack_dir, n))
            Image.fromarray(inv[k]).save(os.path.join(inv_dir, n))
    return len(names)


if __name__ == "__main__":
    main()
