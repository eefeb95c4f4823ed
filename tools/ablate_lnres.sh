#!/bin/bash
# development: time the LayerNorm+residual GEMMs (proj, Mlp.linear2) with parts of the epilogue switched off
for d in 0 4 1 2 3; do
  echo "== PANGU_B200_GEMM_DEBUG=$d (bit0 no residual loads, bit1 no stores, bit2 mainloop only)"
  PANGU_B200_GEMM_DEBUG=$d timeout 200 python tools/kernel_times.py 2>&1 | grep -E "proj_ln_res|mlp_ln_res|mlp1"
done
