mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-unet --no-cpu-baseline > gpurun_out/s32_bench_n8.json 2> gpurun_out/s32_bench_n8.err; echo "bench8 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s32_bench_n8.json'));print(d['n_gpus'],d['value'],d['e2e']['value'],d['clocks'])"
