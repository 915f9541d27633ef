#!/bin/bash
set +e
mkdir -p gpurun_out
sed -n '1,4p;7p;10p;12p;19p' tools/dec2_cases.txt > /tmp/c.txt
CASE_TIMEOUT=30 bash tools/case_sweep.sh < /tmp/c.txt 2>&1 | tee gpurun_out/dec2_sweep.log | grep -v " OK$"
echo "sweep: $(grep -c ' OK$' gpurun_out/dec2_sweep.log) OK of $(wc -l < gpurun_out/dec2_sweep.log)"
timeout -s KILL 100 python tools/microbench.py --ms 1 --shapes q_o,qkv,gate_up,down 2>&1 | grep "M="
