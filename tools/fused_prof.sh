export PANGU_B200_MLP_FUSED=1
timeout 200 python tools/kernel_times.py 2>&1 | grep -E "mlp_ln"
B="python bench.py --steps 1 --warmup 1 --no-cpu"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_fused" --launch-skip 3 -c 1 -f -o gpurun_out/prof_r01c_mlp_fused_lo $B > gpurun_out/prof_fused_lo.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_fused" --launch-skip 0 -c 1 -f -o gpurun_out/prof_r01c_mlp_fused_hi $B > gpurun_out/prof_fused_hi.log 2>&1
ls gpurun_out | grep fused
