"""Golden vectors for the checkpoint-export arithmetic, produced by the UNMODIFIED reference functions
(/root/reference/paroquant/cli/convert.py: _pack_awq, _to_awq_buffers, _quantize_rotated_weight) in this container.
The reference rotates the weight with its CUDA op inside _quantize_rotated_weight; the op is stubbed with the identity
here (theta = 0 semantics) so the integer / rounding arithmetic after it runs on CPU -- the rotation itself is pinned by
tests/golden/ref_gpu_*.npz.

    python tools/gen_convert_golden.py        # writes tests/golden/ref_convert.npz   (needs /root/reference)
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference")
stub = types.ModuleType("paroquant.kernels.cuda")
stub.scaled_pairwise_rotation = lambda x, pairs, theta, scales, group_size: x
sys.modules["paroquant.kernels.cuda"] = stub
import paroquant.kernels  # noqa: E402,F401
sys.modules["paroquant.kernels"].cuda = stub
from paroquant.cli import convert as ref  # noqa: E402

g = torch.Generator().manual_seed(20260924)
N, K, G = 24, 384, 128
weight = torch.randn(N, K, generator=g) * 0.05
scales_flat = (torch.rand(N * K // G, 1, generator=g) * 0.004 + 0.002)
zp_flat = -(torch.rand(N * K // G, 1, generator=g) * 18 - 1.5)          # some clamp at both ends
cs = torch.ones(1, K)
q, s2, z2 = ref._quantize_rotated_weight(weight=weight, pairs=torch.zeros(8, K, dtype=torch.int16), theta=torch.zeros(8, K // 2),
                                         channel_scales=cs, scales_flat=scales_flat, zp_flat=zp_flat, bits=4, group_size=G, device="cpu")
bufs = ref._to_awq_buffers(q, s2, z2)
vals = torch.randint(0, 16, (5, 64), generator=g)
np.savez_compressed(ROOT / "tests" / "golden" / "ref_convert.npz",
                    weight=weight.numpy(), scales_flat=scales_flat.numpy(), zp_flat=zp_flat.numpy(),
                    quantized=q.numpy(), scales_2d=s2.numpy(), zeros_2d=z2.numpy(),
                    qweight=bufs["qweight"].numpy(), qzeros=bufs["qzeros"].numpy(), scales=bufs["scales"].numpy(),
                    pack_in=vals.numpy(), pack_out=ref._pack_awq(vals).numpy())
print("written", ROOT / "tests" / "golden" / "ref_convert.npz", "clamped low/high:", int((q == 0).sum()), int((q == 15).sum()))
