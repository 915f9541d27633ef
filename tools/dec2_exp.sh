#!/bin/bash
set +e
mkdir -p gpurun_out
head -8 tools/dec2_cases.txt > /tmp/c.txt; sed -n '10,11p;19,21p' tools/dec2_cases.txt >> /tmp/c.txt
CASE_TIMEOUT=40 bash tools/case_sweep.sh < /tmp/c.txt 2>&1 | tee gpurun_out/dec2_sweep.log | grep -v " OK$"
echo "sweep: $(grep -c ' OK$' gpurun_out/dec2_sweep.log) OK of $(wc -l < gpurun_out/dec2_sweep.log)"
run() { echo "== $*"; env "$@" timeout -s KILL 200 python tools/microbench.py --ms ${MS:-1} --shapes ${SHAPES:-q_o,qkv,gate_up,down} 2>&1 | grep "M=" ; }
run UNROLL=4
cp paroquant_b200/lib/libparo_b200.so /tmp/main.so
cp paroquant_b200/lib/libparo_b200.u2.so paroquant_b200/lib/libparo_b200.so; run UNROLL=2
cp paroquant_b200/lib/libparo_b200.u1.so paroquant_b200/lib/libparo_b200.so; run UNROLL=1
cp /tmp/main.so paroquant_b200/lib/libparo_b200.so
timeout -s KILL 90 python tools/trace_decode.py gate_up 1 2>&1 | head -12
timeout -s KILL 90 python tools/trace_decode.py q_o 1 2>&1 | head -12
