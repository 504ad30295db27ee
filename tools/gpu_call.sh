mkdir -p gpurun_out
for mode in 1 0; do
AGENDA_Q_CHUNKS=$mode timeout 400 python bench.py --steps 5 --warmup 3 --no-unet --no-cpu-baseline > gpurun_out/s25_bench_$mode.json 2> gpurun_out/s25_bench_$mode.err; python -c "
import json;d=json.load(open('gpurun_out/s25_bench_$mode.json'));k=d['kernels'];print($mode, round(d['value'],2), {n:(round(k[n]['avg_launch_ms']*1000,1), round(k[n]['ms_per_denoise_step_all_layers'],3), round(k[n]['frac'],3)) for n in ('cross_attention_heat','to_q_linear_split')})"
done
