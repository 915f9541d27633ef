#!/bin/bash
# One gpurun call: smoke -> GPU tests -> reference fixtures -> bench -> microbench -> ncu launch list.
# Every step has its own timeout; later steps run even if earlier ones fail.  Logs in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== smoke"; timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
fi
if [ "${GOLDEN:-0}" = "1" ]; then
echo "== golden"; timeout -s KILL 900 python tools/ref_gpu.py golden gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/golden.log
fi
echo "== bench"; timeout -s KILL 900 python bench.py --steps ${STEPS:-200} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "${MICRO:-1}" = "1" ]; then
echo "== microbench"; timeout -s KILL 900 python tools/microbench.py --out gpurun_out/micro.json ${MICRO_ARGS:-} > gpurun_out/micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/micro.log | tail -20
fi
if [ "${NCU:-0}" = "1" ]; then
echo "== ncu launch list"; timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_kernel -s 256 -c 256 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full decode"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 2 -c 1 -o gpurun_out/prof_decode_gate_up python tools/prof_decode.py gate_up 1 4 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full gemm"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 1 -c 1 -o gpurun_out/prof_gemm_gate_up python tools/gemm_bench.py --shapes gate_up --ms 4096 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
echo "== gemm bench"; timeout -s KILL 600 python tools/gemm_bench.py --out gpurun_out/gemm.json > gpurun_out/gemm.log 2>&1; tail -10 gpurun_out/gemm.log
fi
if [ "${REFBENCH:-0}" = "1" ]; then
echo "== reference GPU path"; timeout -s KILL 1200 python tools/ref_gpu.py bench gpurun_out/ref_gpu_bench.json --m 1 > gpurun_out/ref_gpu_bench.log 2>&1; echo "refbench rc=$?"; tail -2 gpurun_out/ref_gpu_bench.log
fi
echo "== done"
