#!/bin/bash
set +e
mkdir -p gpurun_out
for sh in o gate_up; do PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --shape $sh --m 1 2>&1 | tail -21; done
PARO_DECODE_TRACE=1 PARO_NO_PDL=1 timeout -s KILL 200 python tools/stream_trace.py --shape o --m 1 2>&1 | tail -21
PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --chain --m 1 2>&1 | tail -86
echo "== GEMM path for M = 8, 16 (PARO_SMALL_M_MAX=4)"; PARO_SMALL_M_MAX=4 timeout -s KILL 300 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 8,16 2>&1 | tail -8
echo "== done"
