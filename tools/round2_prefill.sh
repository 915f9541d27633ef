#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (rotate + linear + reference)"; timeout -s KILL 1800 python -m pytest tests/test_gpu_rotate.py tests/test_gpu_linear.py tests/test_gpu_reference.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== rotate: ours vs reference"; timeout -s KILL 300 python tools/ref_gpu.py rotbench ours 2>&1 | tail -8; timeout -s KILL 300 python tools/ref_gpu.py rotbench reference 2>&1 | tail -8
echo "== gemm bench"; timeout -s KILL 600 python tools/gemm_bench.py --out gpurun_out/r02_gemm.json 2>&1 | tail -14
echo "== done"
