#!/bin/bash
# tensor-parallel chain on 2 GPUs: parity (small, then Llama shapes) and timing against linears + NCCL
set +e
mkdir -p gpurun_out
N=${N:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== tp small m=1"; timeout -s KILL 300 $RUN tools/tp_check.py --small --batch 1 2>&1 | grep -v "^W0\|^\*\*\*" | tail -6
echo "== tp small m=4"; timeout -s KILL 300 $RUN tools/tp_check.py --small --batch 4 2>&1 | grep -v "^W0\|^\*\*\*" | tail -6
echo "== tp llama m=1"; timeout -s KILL 600 $RUN tools/tp_check.py --batch 1 2>&1 | grep -v "^W0\|^\*\*\*" | tail -8
echo "== done"
