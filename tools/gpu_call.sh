mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/s21_pytest.log; tail -2 gpurun_out/s21_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s21_smoke.log 2>&1; tail -1 gpurun_out/s21_smoke.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s21_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['full_unet_step']['value'],{k:round(v.get('frac',0),3) for k,v in d['kernels'].items()}, d['kernels']['cross_attention_backward']['ms'])"
