"""Tiny driver for ncu: a few back-to-back launches of the fused small-M kernel on distinct weight sets.
    ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 2 -o gpurun_out/prof python tools/prof_decode.py q_o 1
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from paroquant_b200 import _cabi  # noqa: E402
from paroquant_b200.checkpoint import make_synthetic_activations, make_synthetic_layer  # noqa: E402
from paroquant_b200.linear import ParoLinearKernel  # noqa: E402
from tools.microbench import SHAPES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "q_o"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nsets = int(sys.argv[3]) if len(sys.argv) > 3 else 8
K, parts = SHAPES[name]
ks = [ParoLinearKernel.from_buffers(make_synthetic_layer(K, parts, seed=900 + i, device="cuda"), torch.bfloat16,
                                    check_pairs=False, max_m=M) for i in range(nsets)]
x = make_synthetic_activations(M, K, seed=1, device="cuda")
y = torch.empty(M, sum(parts), dtype=torch.bfloat16, device="cuda")
torch.cuda.synchronize()
for k in ks:
    _cabi.linear_forward(k.shape, k.packed, x, None, k.workspace, out=y)
torch.cuda.synchronize()
print("done")
