#!/bin/bash
set +e
mkdir -p gpurun_out
cat tools/dec2_cases.txt > /tmp/c.txt
cat >> /tmp/c.txt <<'EOT'
4096 4096,1024,1024 16 bfloat16
4096 4096 12 float16
11008 4096 6 bfloat16
EOT
CASE_TIMEOUT=40 bash tools/case_sweep.sh < /tmp/c.txt 2>&1 | tee gpurun_out/dec2_sweep.log | grep -v " OK$"
echo "sweep: $(grep -c ' OK$' gpurun_out/dec2_sweep.log) OK of $(wc -l < gpurun_out/dec2_sweep.log)"
timeout -s KILL 200 python tools/microbench.py --ms 4,16 --shapes q_o,qkv,gate_up,down 2>&1 | grep "M="
