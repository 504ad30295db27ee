mkdir -p gpurun_out
timeout 300 python tools/profile_unet.py > gpurun_out/s16_unet_profile.txt 2>&1; head -36 gpurun_out/s16_unet_profile.txt | cut -c1-60,150-215
timeout 400 python - <<'PY' 2>&1 | tail -2
import sys, json
sys.argv = ["bench.py"]
import torch, bench
out = bench.full_unet_step(torch.device("cuda", 0), 8, reps=2)
print(out["value"], out["ms_per_denoise_step"])
PY
