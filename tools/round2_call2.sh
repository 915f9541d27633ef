#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== pytest gpu B8 off"; PARO_DECODE_B8=0 timeout -s KILL 900 python -m pytest tests/test_gpu_linear.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_b16.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_b16.log
for sh in o gate_up; do PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --shape $sh --m 1 2>&1 | tail -14; done
PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --shape gate_up --m 16 2>&1 | tail -14
PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --chain --m 1 2>&1 | tail -50
echo "== chain bench"; timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -4
for c in 1 4 8; do echo "== microbench C=$c"; PARO_DECODE_C=$c timeout -s KILL 300 python tools/microbench.py --shapes q_o,gate_up,down --ms 1 2>&1 | tail -3; done
echo "== microbench sets 4 / 6"; for s in 4 6; do PARO_DECODE_SETS=$s timeout -s KILL 300 python tools/microbench.py --shapes q_o,gate_up --ms 1 2>&1 | tail -2; done
echo "== done"
