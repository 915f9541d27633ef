#!/bin/bash
# decode prologue restructure (rotation before the CTA sync, last warps rotate, no zero fill): parity, then timings
set +e
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -s KILL 420 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== microbench sets=5"; timeout -s KILL 200 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1,4,16 --out gpurun_out/mb_new.json 2>&1 | tail -12
echo "== microbench sets=6"; PARO_DECODE_SETS=6 timeout -s KILL 100 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1 2>&1 | tail -4
echo "== microbench sets=7"; PARO_DECODE_SETS=7 timeout -s KILL 100 python tools/microbench.py --shapes gate_up,down --ms 1 2>&1 | tail -2
echo "== trace"; for s in q_o gate_up; do timeout 100 python tools/trace_decode.py $s 1 2>&1 | grep -v "^    "; done
echo "== done"
