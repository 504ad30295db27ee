mkdir -p gpurun_out
export AGENDA_KNOBS=1
timeout 600 python -m pytest tests/test_gpu_post.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/s6_pytest.log; tail -3 gpurun_out/s6_pytest.log
run() { echo "== $*"; env "$@" timeout 60 python tools/bench_ccl.py 2048 512 2>&1 | head -2; }
(run X=0; run AGENDA_CCL_TAB=0; run AGENDA_CCL_HINTS=-1) 2>&1 | tee gpurun_out/s6_ccl_sweep.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ccl_bbox_cta -s 1 -c 1 -o gpurun_out/s6_ccl_full -f python tools/probe_kernels.py ccl > gpurun_out/s6_ncu_ccl.log 2>&1; echo "ncu rc=$?"
