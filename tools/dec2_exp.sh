#!/bin/bash
set +e
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout -s KILL 200 python tools/microbench.py --ms ${MS:-1} --shapes ${SHAPES:-gate_up,down} 2>&1 | grep "M=" ; }
run PARO_DECODE_SETS=6
run PARO_DECODE_SETS=6 PARO_DECODE_FAKE1=1
run PARO_DECODE_SETS=6 PARO_DECODE_STAGES=8
run PARO_DECODE_SETS=6 PARO_NO_PDL=1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 2 -c 1 -o gpurun_out/prof_dec2_gate_up python tools/prof_decode.py gate_up 1 4 > gpurun_out/ncu_dec2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dec2.log
