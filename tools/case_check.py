"""Run one (K, parts, M, dtype) case of the fused linear against a torch reference built from the
kernel's own dequantised operand + the standalone rotate kernel.  Used with a process timeout to
isolate hangs:   python tools/case_check.py 4096 4096 1 bfloat16
"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import paroquant_b200.kernels.cuda  # noqa
from paroquant_b200.checkpoint import make_synthetic_activations, make_synthetic_layer
from paroquant_b200.linear import ParoLinearKernel

K = int(sys.argv[1]); parts = [int(v) for v in sys.argv[2].split(",")]; M = int(sys.argv[3])
dt = getattr(torch, sys.argv[4]) if len(sys.argv) > 4 else torch.bfloat16
L = make_synthetic_layer(K, parts, seed=47, device="cuda")
k = ParoLinearKernel.from_buffers(L, dt, max_m=M)
W = k.dense_weight().double()
x = make_synthetic_activations(M, K, seed=60 + M, device="cuda", dtype=dt)
y = k(x)
torch.cuda.synchronize()
n0, chunks = 0, []
for p, n in enumerate(parts):
    xr = torch.ops.rotation.rotate(x, L.pairs[p], L.theta[p], L.channel_scales[p]).double()
    chunks.append(xr @ W[:, n0:n0 + n]); n0 += n
ref = torch.cat(chunks, -1).float().to(dt).double()
err = ((y.double() - ref).norm() / ref.norm()).item()
y2 = k(x); torch.cuda.synchronize()
print(f"K={K} parts={parts} M={M} {sys.argv[4] if len(sys.argv)>4 else 'bfloat16'}: rel_err={err:.3e} deterministic={torch.equal(y, y2)} {'OK' if err < 3e-4 else 'BAD'}")
