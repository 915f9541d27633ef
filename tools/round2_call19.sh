#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (linear, reference)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_linear.py tests/test_gpu_reference.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== microbench (bulk-copied pre-rotated x from M = 8)"; timeout -s KILL 600 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 8,16 2>&1 | tail -8
echo "== microbench (from M = 4)"; PARO_DECODE_PREROT_M=4 timeout -s KILL 600 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 4,5 2>&1 | tail -8
echo "== done"
