#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== ncu launch list of the bench command"; timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stream_kernel|decode_kernel" -c 1400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-prefill --no-ref-gpu > gpurun_out/r02_ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-300; wc -l gpurun_out/r02_launches.csv
echo "== ncu full: chain"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_chain python tools/prof_chain.py > gpurun_out/r02_ncu_chain.log 2>&1; echo "rc=$?"
for sh in q_o qkv gate_up; do echo "== ncu full: decode_kernel $sh"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_decode_$sh python tools/prof_decode.py $sh 1 6 > gpurun_out/r02_ncu_$sh.log 2>&1; echo "rc=$?"; done
echo "== bench (with neighbour kernels)"; timeout -s KILL 900 python bench.py --steps 20 --warmup 5 --no-prefill --no-ref-gpu --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('value', d['value'], 'chain', d['chain']['tokens_per_s'], 'per_linear', d['per_linear']['tokens_per_s'], 'with neighbours', d['per_linear']['with_neighbour_kernels']['tokens_per_s'])"
ls -la gpurun_out/*.ncu-rep
echo "== done"
