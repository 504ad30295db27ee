mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_self_bwd.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/s9_train.json 2> gpurun_out/s9_train.err; echo "train rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s9_train.json'));print(d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['frac'])"
