#!/bin/bash
# correctness sweep (one process per case, short timeouts) -> microbench -> phase trace of the small-M kernel
set +e
mkdir -p gpurun_out
if [ "${SWEEP:-1}" = "1" ]; then
  CASE_TIMEOUT=${CASE_TIMEOUT:-40} bash tools/case_sweep.sh < ${CASES:-tools/dec2_cases.txt} 2>&1 | tee gpurun_out/dec2_sweep.log
fi
for sets in ${SETS_LIST:-6}; do
  echo "== microbench SETS=$sets"
  PARO_DECODE_SETS=$sets timeout -s KILL 200 python tools/microbench.py --ms ${MS:-1} --shapes ${SHAPES:-q_o,qkv,gate_up,down} --out gpurun_out/micro_s$sets.json 2>&1 | tail -8
done
if [ "${TRACE:-1}" = "1" ]; then
  for sh in ${TRACE_SHAPES:-gate_up q_o}; do
    timeout -s KILL 90 python tools/trace_decode.py $sh 1 2>&1 | tail -12
  done
fi
if [ "${BENCH:-0}" = "1" ]; then
  timeout -s KILL 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-prefill 2>/dev/null | cut -c1-400
fi
