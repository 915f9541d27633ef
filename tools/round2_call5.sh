#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (chain + linear)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_chain.py tests/test_gpu_linear.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for sh in o gate_up; do PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --shape $sh --m 1 2>&1 | tail -12; done
PARO_DECODE_TRACE=1 timeout -s KILL 200 python tools/stream_trace.py --chain --m 1 2>&1 | tail -50
echo "== chain bench"; timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -4
echo "== microbench"; timeout -s KILL 600 python tools/microbench.py --out gpurun_out/micro_r2_stream.json --shapes q_o,qkv,gate_up,down > gpurun_out/micro_r2_stream.log 2>&1; tail -14 gpurun_out/micro_r2_stream.log
for f in 3 12 23; do echo "== inflight $f"; PARO_DECODE_INFLIGHT=$f timeout -s KILL 300 python tools/microbench.py --shapes q_o,gate_up,down --ms 1 2>&1 | tail -3; PARO_DECODE_INFLIGHT=$f timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2; done
echo "== microbench C=4"; PARO_DECODE_C=4 timeout -s KILL 300 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1,16 2>&1 | tail -8
echo "== done"
