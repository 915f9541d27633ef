#!/bin/bash
set +e
mkdir -p gpurun_out
CASE_TIMEOUT=40 bash tools/case_sweep.sh < tools/dec2_cases.txt 2>&1 | tee gpurun_out/dec2_sweep.log | grep -v " OK$"
echo "sweep: $(grep -c ' OK$' gpurun_out/dec2_sweep.log) OK of $(wc -l < gpurun_out/dec2_sweep.log)"
CASE_TIMEOUT=40 bash tools/case_sweep.sh < tools/gemm_bisect.txt 2>&1 | tail -6
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -5
run() { echo "== $*"; env "$@" timeout -s KILL 200 python tools/microbench.py --ms ${MS:-1} --shapes ${SHAPES:-q_o,qkv,gate_up,down} 2>&1 | grep "M=" ; }
run PARO_DECODE_SETS=6
run PARO_DECODE_SETS=5
MS=16 run PARO_DECODE_SETS=6
timeout -s KILL 300 python tools/gemm_bench.py 2>&1 | tail -10
timeout -s KILL 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-prefill 2>/dev/null | cut -c1-600
