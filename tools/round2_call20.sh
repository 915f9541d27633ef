#!/bin/bash
# group_size 64 + fused rotate backward: full GPU suite at HEAD, then a quick bench for regressions of the G = 128 path
set +e
mkdir -p gpurun_out
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu (all)"; timeout -s KILL 420 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench quick"; timeout -s KILL 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
echo "== done"
