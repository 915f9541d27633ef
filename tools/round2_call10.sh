#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (chain + linear)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_chain.py tests/test_gpu_linear.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --chain --m 1 2>&1 | tail -86 | grep -v "rows (slowest"
echo "== chain bench"; timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -4
echo "== chain bench C=4"; PARO_DECODE_C=4 timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2
echo "== chain bench m=4"; timeout -s KILL 300 python tools/chain_bench.py --m 4 2>&1 | tail -2
echo "== microbench"; timeout -s KILL 600 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1 2>&1 | tail -4
echo "== microbench V1"; PARO_DECODE_V1=1 timeout -s KILL 600 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1,4,16 2>&1 | tail -12
echo "== done"
