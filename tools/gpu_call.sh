mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-unet > gpurun_out/s13_bench_n2.json 2> gpurun_out/s13_bench_n2.err; echo "bench2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s13_bench_n2.json'));print(d['n_gpus'],d['value'],d['e2e'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference > gpurun_out/s13_bench_ref_n2.json 2> gpurun_out/s13_bench_ref_n2.err; echo "ref2 rc=$?"; cut -c1-200 gpurun_out/s13_bench_ref_n2.json
