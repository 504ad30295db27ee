#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list, ncu --set full of top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
python tools/bench_attn.py 16 4096 8 40 0 > gpurun_out/bench_attn.log 2>&1
python tools/bench_attn.py 32 9216 5 64 0 >> gpurun_out/bench_attn.log 2>&1
python tools/bench_ccl.py > gpurun_out/bench_ccl.log 2>&1
cat gpurun_out/bench_attn.log gpurun_out/bench_ccl.log
# ncu launch list of the bench command (cold-cache, serialised: shares matter, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full-set captures of the top kernels (eager replay so every launch is visible)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_self_sm100_v2 -s 0 -c 1 \
  -o gpurun_out/self_attn_full -f python tools/run_steps.py 1 8 > gpurun_out/ncu_self.log 2>&1; echo "ncu self rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_cross_sm100_res -s 0 -c 1 \
  -o gpurun_out/cross_attn_full -f python tools/run_steps.py 1 8 > gpurun_out/ncu_cross.log 2>&1; echo "ncu cross rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ccl_bbox_cta -s 1 -c 1 \
  -o gpurun_out/ccl_full -f python tools/bench_ccl.py > gpurun_out/ncu_ccl.log 2>&1; echo "ncu ccl rc=$?"
ls -la gpurun_out
