#!/bin/bash
# Scratch script of a `gpurun` visit (rewritten per call); this version is the full regression of the round:
# GPU parity tests, smoke(), the bench line with the reference arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.log; tail -2 gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
