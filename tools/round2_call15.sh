#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (experts, checkpoint, chain)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_experts.py tests/test_gpu_checkpoint.py tests/test_gpu_chain.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== ncu: prefill cells"; 
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/r02_prof_gemm_o_m256 python tools/gemm_bench.py --shapes q_o --ms 256 > gpurun_out/r02_ncu_gemm1.log 2>&1; echo "rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/r02_prof_gemm_qkv_m4096 python tools/gemm_bench.py --shapes qkv --ms 4096 > gpurun_out/r02_ncu_gemm2.log 2>&1; echo "rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:rotate_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_rotate_m4096 python tools/gemm_bench.py --shapes q_o --ms 4096 > gpurun_out/r02_ncu_rot.log 2>&1; echo "rc=$?"
echo "== ncu launch list: qkv prefill 4096"; timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_prefill.csv python tools/gemm_bench.py --shapes qkv,q_o,down --ms 4096 > /dev/null 2>&1; tail -25 gpurun_out/r02_launches_prefill.csv | cut -d, -f5,13-15 | tail -22
echo "== done"
