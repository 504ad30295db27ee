mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heat_upsample_accum_tiled -s 1 -c 1 -o gpurun_out/s23_up_full -f python tools/probe_kernels.py upsample > gpurun_out/s23_ncu_up.log 2>&1; echo "ncu rc=$?"
