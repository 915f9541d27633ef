#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== parity subset"; timeout -s KILL 200 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k "fused_linear_vs_oracle or llama_shapes" --timeout 120 -p no:cacheprovider 2>&1 | tail -4
echo "== microbench"; timeout -s KILL 200 python tools/microbench.py --shapes q_o,qkv,gate_up,down --ms 1,4,16 --out gpurun_out/mb_new2.json 2>&1 | tail -12
echo "== trace"; for s in q_o down; do timeout 100 python tools/trace_decode.py $s 1 2>&1 | grep -v "^    "; done
echo "== done"
