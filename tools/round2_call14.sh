#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (linear)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_linear.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== gemm bench"; timeout -s KILL 600 python tools/gemm_bench.py --out gpurun_out/r02_gemm.json --shapes q_o,qkv,gate_up,down,l2_down 2>&1 | tail -16
echo "== done"
