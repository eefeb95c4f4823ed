export PANGU_B200_MLP_FUSED=1
B="python bench.py --steps 1 --warmup 1 --no-cpu"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_fused" --launch-skip 0 -c 1 -f -o gpurun_out/prof_r01e_mlp_fused_hi $B > gpurun_out/prof_fused_hi.log 2>&1
ls -la gpurun_out | grep r01e
