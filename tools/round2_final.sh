#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== smoke"; timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu (all)"; timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout -s KILL 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== gemm bench"; timeout -s KILL 600 python tools/gemm_bench.py --out gpurun_out/r02_gemm.json --shapes q_o,qkv,gate_up,down,l2_down 2>&1 | tail -16
echo "== done"
