#!/bin/bash
# round 2, call 1: GPU tests with the new small-M kernel (paro_stream.cu), A/B microbench against round 1's kernel, launch floor
set +e
mkdir -p gpurun_out
echo "== smoke"; timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (stream kernel)"; timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== pytest gpu B8 off"; PARO_DECODE_B8=0 timeout -s KILL 900 python -m pytest tests/test_gpu_linear.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_b16.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_b16.log
echo "== pytest gpu V1"; PARO_DECODE_V1=1 timeout -s KILL 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_reference.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_v1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_v1.log
echo "== launch probe"; timeout -s KILL 120 paroquant_b200/lib/launch_probe > gpurun_out/launch_probe.log 2>&1; cat gpurun_out/launch_probe.log
echo "== microbench stream"; timeout -s KILL 600 python tools/microbench.py --out gpurun_out/micro_r2_stream.json --shapes q_o,qkv,gate_up,down > gpurun_out/micro_r2_stream.log 2>&1; tail -14 gpurun_out/micro_r2_stream.log
echo "== microbench V1"; PARO_DECODE_V1=1 timeout -s KILL 600 python tools/microbench.py --out gpurun_out/micro_r2_v1.json --shapes q_o,qkv,gate_up,down --ms 1 > gpurun_out/micro_r2_v1.log 2>&1; tail -6 gpurun_out/micro_r2_v1.log
echo "== done"
