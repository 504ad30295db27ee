"""GPU, world size 2: BASELINE configs[2] through bench.py --workload config3 — seeds sharded over two B200s, records
gathered over NCCL (agenda_b200.sharding.gather_records), compared bit for bit with a one-GPU run of the same seeds.
Skipped on a box with fewer than two GPUs (the driver's one-GPU test run)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_bench(n_gpus, dump, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(n_gpus), "--workload",
           "config3", "--num-images", "21", "--images-per-step", "4", "--denoise-steps", "2", "--steps", "1", "--warmup",
           "1", "--dump", dump]
    if n_gpus == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py")] + cmd[cmd.index("--gpus"):]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_config3_two_gpus_equal_one_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    two = _run_bench(2, str(tmp_path / "two.npz"), 29611)
    one = _run_bench(1, str(tmp_path / "one.npz"), 29612)
    assert two["n_gpus"] == 2 and one["n_gpus"] == 1 and two["scaling"] == "strong"
    assert two["determinism_check"]["bit_identical"] and one["determinism_check"]["bit_identical"]
    assert two["gathered"]["images"] == 21                      # ragged: rank 0 has 11 seeds, rank 1 has 10
    a, b = np.load(tmp_path / "two.npz"), np.load(tmp_path / "one.npz")
    for k in ("heat", "stack", "counts", "boxes"):
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
