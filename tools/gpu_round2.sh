#!/bin/bash
# Round-2 GPU visit: tests, bench, launch list, ncu full captures of the top kernels.  Usage: bash tools/gpu_round2.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
if [ "$2" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for spec in "self:attn_self_sm100_v2:run_steps.py 1 8" "x3:attn_cross_sm100_x3:run_steps.py 1 8" "up:heat_upsample_accum_tiled:probe_kernels.py upsample" "post:postprocess_stack:probe_kernels.py postprocess" "xbwd:attn_cross_bwd:probe_kernels.py cross_bwd" "ccl:ccl_bbox_cta:probe_kernels.py ccl"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%%:*}; cmd=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -o gpurun_out/${TAG}_${name}_full -f python tools/$cmd > gpurun_out/${TAG}_ncu_${name}.log 2>&1; echo "ncu $name rc=$?"
done
fi
