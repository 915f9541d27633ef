"""A few launches of the fused rotate backward (fp32, 4096 x 4096, krot 8) for an ncu capture."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paroquant_b200.kernels.cuda  # noqa: F401,E402
from paroquant_b200 import _cabi  # noqa: E402
from paroquant_b200.checkpoint import make_synthetic_layer  # noqa: E402

M, K, G = 4096, 4096, 128
L = make_synthetic_layer(K, [64], seed=5, device="cuda")
pr, th, sc = L.pairs[0], L.theta[0].float(), L.channel_scales[0].float().view(-1)
x = torch.randn(M, K, device="cuda")
go = torch.randn(M, K, device="cuda")
y = torch.ops.rotation.rotate(x, pr, th, sc, G)
for _ in range(3):
    _cabi.rotate_backward(y, go, x, pr, th, sc, G)
torch.cuda.synchronize()
