#!/bin/bash
# producer: how many ring stages before the dependency wait (PARO_DECODE_EARLY; unset = all, the shipped behaviour)
set +e
mkdir -p gpurun_out
for e in all 0 2 4 8 12 all; do
  echo "== early=$e"; if [ $e = all ]; then unset PARO_DECODE_EARLY; else export PARO_DECODE_EARLY=$e; fi
  timeout -s KILL 100 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1 2>&1 | tail -4
done 2>&1 | tee gpurun_out/ab_early.txt
echo "== done"
