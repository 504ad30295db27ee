"""In-tree build of libagenda_b200.so: nvcc, sm_100a only, -lineinfo (ncu source pages), static cudart.

    python -m agenda_b200.build            # build if sources are newer than the library
    python -m agenda_b200.build --force
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libagenda_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-cudart", "static", "--expt-relaxed-constexpr", "--threads", "0"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def _variants_wanted() -> bool:
    return os.environ.get("AGENDA_BUILD_VARIANTS", "0") == "1"


def build(force: bool = False, verbose: bool = False, variants=None) -> str:
    """variants=True (or AGENDA_BUILD_VARIANTS=1) adds -DAGENDA_VARIANTS: the measurement / test variants of the
    self-attention kernel (agenda_attn_self_fwd_variant 10..58).  The product library does not carry them."""
    variants = _variants_wanted() if variants is None else variants
    flavor = "variants" if variants else "product"
    stamp = LIB + ".flavor"   # which of the two the library on disk is: asking for the other one rebuilds
    have = open(stamp).read().strip() if os.path.exists(stamp) else "product"
    if not force and not _stale() and have == flavor:
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libagenda_b200.so")
    cmd = ([nvcc] + NVCC_FLAGS + (["-DAGENDA_VARIANTS"] if variants else []) + (["-Xptxas", "-v"] if verbose else [])
           + ["-o", LIB] + sources())
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(stamp, "w") as f:
        f.write(flavor + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variants=True if "--variants" in sys.argv else None))
