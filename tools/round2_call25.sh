#!/bin/bash
# end-of-round records: full bench line at N = 1, training-op bench, group_size 64 microbench
set +e
mkdir -p gpurun_out
echo "== bench N=1"; timeout -s KILL 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
echo "== backward"; timeout -s KILL 90 python tools/backward_bench.py --out gpurun_out/backward_bench.json 2>&1 | tail -6
echo "== group 64"; timeout -s KILL 90 python tools/microbench.py --group 64 --shapes q_o,qkv,gate_up,down --ms 1,16 --out gpurun_out/mb_group64.json 2>&1 | tail -8
echo "== done"
