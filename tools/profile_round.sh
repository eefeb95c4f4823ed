#!/bin/bash
# Round profile pass on the GPU box: launch list + one ncu --set full capture per hot kernel.
# usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/)
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.log 2>&1
cap() {  # name, kernel regex, launches to skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $B > gpurun_out/prof_${TAG}_$1.log 2>&1
}
# per forward: QKV hi x2, lo x12, hi x2; same order for attention / LN GEMMs; Mlp.linear1 (CfgMLP1) only at lo
cap mlp_fused2_lo mlp_fused2_kernel 2
cap CfgLNRes384_proj CfgLNRes384 2
cap attention_lo window_attention_tc 4
cap attention_hi window_attention_tc 0
cap CfgQKV_lo CfgQKV 4
cap CfgQKV_hi CfgQKV 0
cap mlp_fused_hi mlp_fused_kernel 0
ls -la gpurun_out | tail -12
