mkdir -p gpurun_out
export AGENDA_KNOBS=1
timeout 600 python -m pytest tests/test_gpu_post.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/s4_pytest.log; tail -3 gpurun_out/s4_pytest.log
run() { echo "== $*"; env "$@" timeout 60 python tools/bench_ccl.py 2048 512 2>&1 | head -2; }
(run X=0; run AGENDA_CCL_TAB=0; run AGENDA_CCL_HINTS=-1; run AGENDA_CCL_HINTS=1; run AGENDA_CCL_HINTS=25;
 run AGENDA_CCL_CTA_SMEM_KB=100; run AGENDA_CCL_CTA_THREADS=512 AGENDA_CCL_CTA_SMEM_KB=56; run AGENDA_CCL_CTA_THREADS=512 AGENDA_CCL_CTA_SMEM_KB=72) > gpurun_out/s4_ccl_sweep.txt 2>&1
cat gpurun_out/s4_ccl_sweep.txt
unset AGENDA_KNOBS
timeout 400 python bench.py --steps 5 --warmup 3 --no-unet --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/s4_bench.json'));print(d['value'],d['e2e'],d['kernels']['ccl_bbox_512'])"
timeout 400 python bench.py --steps 5 --warmup 3 --no-unet --no-cpu-baseline --e2e-serial > gpurun_out/s4_bench_serial.json 2> gpurun_out/s4_bench_serial.err; python -c "
import json;d=json.load(open('gpurun_out/s4_bench_serial.json'));print(d['value'],d['e2e'])"
