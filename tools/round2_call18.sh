#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (linear, reference)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_linear.py tests/test_gpu_reference.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== microbench (pre-rotation from M = 4)"; timeout -s KILL 600 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1,4,8,16 2>&1 | tail -16
echo "== microbench (in-kernel rotation at every M)"; PARO_DECODE_PREROT_M=17 timeout -s KILL 600 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 2,4,8,16 2>&1 | tail -16
echo "== microbench (pre-rotation from M = 2)"; PARO_DECODE_PREROT_M=2 timeout -s KILL 600 python tools/microbench.py --shapes q_o,down --ms 2,3 2>&1 | tail -4
echo "== done"
