#!/bin/bash
# Round profile pass on the GPU box: launch list + one ncu --set full capture per hot kernel.
# usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.log 2>&1
cap() {  # name, kernel regex, launches to skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $B > gpurun_out/prof_${TAG}_$1.log 2>&1
}
cap CfgMLP1_lo CfgMLP1 2
cap CfgLNRes384_mlp2 CfgLNRes384 1
cap attention_lo window_attention_tc 2
cap CfgQKV_lo CfgQKV 2
cap CfgLNRes192_mlp2 CfgLNRes192 1
ls -la gpurun_out
# backward kernels (training step workload)
T="python tools/train_times.py"
capt() {
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $T > gpurun_out/prof_${TAG}_$1.log 2>&1
}
capt attention_bwd_lo window_attention_bwd 6
capt wgrad_lo wgrad_kernel 20
ls -la gpurun_out
