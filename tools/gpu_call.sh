mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu -x 2>&1 | tail -3
timeout 400 python - <<'PY' 2>&1 | tail -3
import sys, json
sys.argv = ["bench.py"]
import torch, bench
out = bench.full_unet_step(torch.device("cuda", 0), 8, reps=2)
print(json.dumps(out))
PY
timeout 300 python tools/profile_unet.py > gpurun_out/s11_unet_profile.txt 2>&1; head -40 gpurun_out/s11_unet_profile.txt | cut -c1-60,150-215
