"""Tiny driver for ncu: a few decoder-tail chain launches (o -> gate_up -> down -> qkv, M = 1) on distinct weight sets.
    ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 3 -c 1 -o gpurun_out/prof_chain python tools/prof_chain.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from paroquant_b200 import chain  # noqa: E402
from paroquant_b200.checkpoint import make_synthetic_activations, make_synthetic_layer  # noqa: E402
from paroquant_b200.linear import ParoLinearKernel  # noqa: E402

H, KV, I = 4096, 1024, 14336
SH = {"o": (H, [H]), "gate_up": (H, [I, I]), "down": (I, [H]), "qkv": (H, [H, KV, KV])}
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dt, dev = torch.bfloat16, "cuda"
layers = [{n: ParoLinearKernel.from_buffers(make_synthetic_layer(K, p, seed=500 + 8 * li + i, device=dev), dt, check_pairs=False, max_m=M)
           for i, (n, (K, p)) in enumerate(SH.items())} for li in range(4)]
attn = make_synthetic_activations(M, H, seed=1, device=dev, dtype=dt)
resid = make_synthetic_activations(M, H, seed=2, device=dev, dtype=dt)
w = torch.ones(H, dtype=dt, device=dev)
chains = [chain.decoder_tail(l["o"], l["gate_up"], l["down"], l["qkv"], attn_out=attn, residual=resid, post_attn_norm=w, next_input_norm=w)[0]
          for l in layers]
torch.cuda.synchronize()
for c in chains:
    c()
torch.cuda.synchronize()
print("done")
