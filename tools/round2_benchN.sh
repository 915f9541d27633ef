#!/bin/bash
# the driver's multi-GPU bench command, N ranks on one box
set +e
mkdir -p gpurun_out
N=${N:-2}
echo "== bench N=$N (chains, fused tensor-parallel sum)"; timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-1800; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
echo "== bench N=$N (round-1 path: one launch per linear + NCCL)"; BENCH_TP_NCCL=1 timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?"; cat gpurun_out/bench_n${N}_nccl.json | cut -c1-700
echo "== bench N=$N reference arm"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | cut -c1-300
echo "== done"
