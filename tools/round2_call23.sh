#!/bin/bash
# same-box A/B of decode kernel variants (old prologue vs rotation-before-sync, with / without zero fill, first / last warps rotate)
set +e
mkdir -p gpurun_out
for v in old new zero first firstzero old new; do
  echo "== $v"; PARO_B200_LIB=$PWD/paroquant_b200/lib_variants/libparo_$v.so timeout -s KILL 120 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1,4 2>&1 | tail -8
done 2>&1 | tee gpurun_out/ab_variants.txt
echo "== done"
