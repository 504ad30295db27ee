mkdir -p gpurun_out
timeout 500 python bench.py --workload sd21 --steps 2 --warmup 3 --no-unet --no-cpu-baseline > gpurun_out/s29_sd21.json 2> gpurun_out/s29_sd21.err; echo "sd21 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s29_sd21.json'));print(d['value'],d['roofline']['frac'],d['heat_max_abs_err']['value'], d['kernels']['cross_attention_heat']['frac'])"
timeout 500 python bench.py --workload config3 --num-images 64 --denoise-steps 10 > gpurun_out/s29_c3.json 2> gpurun_out/s29_c3.err; echo "config3 rc=$?"; cut -c1-600 gpurun_out/s29_c3.json; tail -2 gpurun_out/s29_c3.err
