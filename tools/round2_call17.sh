#!/bin/bash
set +e
mkdir -p gpurun_out
for sh in down gate_up q_o; do PARO_DECODE_TRACE=1 PARO_DECODE_VERBOSE=1 timeout -s KILL 200 python tools/trace_decode.py $sh 16 2>&1 | grep -v "^    " | tail -14; done
PARO_DECODE_TRACE=1 PARO_DECODE_VERBOSE=1 timeout -s KILL 200 python tools/trace_decode.py down 4 2>&1 | grep -v "^    " | tail -14
echo "== chain bench (reverted accesses)"; timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2
echo "== done"
