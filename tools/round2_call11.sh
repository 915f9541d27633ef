#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (chain)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_chain.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== chain bench"; timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2
echo "== chain bench C=4"; PARO_DECODE_C=4 timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2
echo "== chain bench C=3"; PARO_DECODE_C=3 timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2
echo "== done"
