#!/bin/bash
set +e
mkdir -p gpurun_out
N=${N:-8}
echo "== bench N=$N (chains, fused tensor-parallel sum)"; timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-400; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -8
echo "== done"
