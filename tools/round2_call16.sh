#!/bin/bash
set +e
mkdir -p gpurun_out
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== pytest gpu (chain, experts, checkpoint)"; timeout -s KILL 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_experts.py tests/test_gpu_checkpoint.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== tp llama m=1"; timeout -s KILL 600 $RUN tools/tp_check.py --batch 1 2>&1 | grep -v "^W0\|^\*\*\*\|OMP\|^$" | tail -5
echo "== tp small m=4"; timeout -s KILL 300 $RUN tools/tp_check.py --small --batch 4 2>&1 | grep -v "^W0\|^\*\*\*\|OMP\|^$" | tail -3
echo "== chain bench"; timeout -s KILL 300 python tools/chain_bench.py --m 1 2>&1 | tail -2
echo "== done"
