mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_self_bwd.py -q -m gpu -x -k "backward or training" 2>&1 | tail -3
timeout 120 python tools/profile_cross_bwd.py 2>&1 | cut -c1-92,180-240 | tail -12 | head -6
timeout 120 python tools/profile_cross_bwd.py 2 1024 8 80 3 2>&1 | cut -c1-92,180-240 | tail -12 | head -6
timeout 300 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/s20_train.json 2> gpurun_out/s20_train.err; echo "train rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s20_train.json'));print(d['ms_per_step'],d['cuda_graph'])"
