mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_self_bwd_kernel|attn_bwd_delta" -s 4 -c 4 -o gpurun_out/s31_self_bwd_full -f python tools/probe_kernels.py self_bwd > gpurun_out/s31_ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_cross_bwd_dq_kernel|attn_self_bwd_kernel" -s 2 -c 2 -o gpurun_out/s31_cross_bwd_full -f python tools/probe_kernels.py cross_bwd_tc > gpurun_out/s31_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"gn_stats|gn_apply|layernorm_kernel|geglu_kernel" -s 4 -c 4 -o gpurun_out/s31_glue_full -f python tools/probe_kernels.py glue > gpurun_out/s31_ncu_c.log 2>&1; echo "ncu c rc=$?"
