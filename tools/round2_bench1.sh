#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== bench N=1"; timeout -s KILL 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout -s KILL 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "== done"
