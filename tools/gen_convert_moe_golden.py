"""Golden vectors for the MoE export arithmetic and checkpoint naming, produced by the UNMODIFIED reference functions
(/root/reference/paroquant/cli/convert.py: _quantize_moe, _inject_quantized_moe_state_dict) in this container.  As in
gen_convert_golden.py the reference's CUDA rotation is stubbed with the identity so everything after it runs on CPU.

    python tools/gen_convert_moe_golden.py        # writes tests/golden/ref_convert_moe.npz   (needs /root/reference)
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

g = torch.Generator().manual_seed(20260925)
E, H, I, G, R = 2, 128, 128, 128, 8
st = {"n_bits": torch.tensor(4), "group_size": torch.tensor(G),
      "gate_up_weight": torch.randn(E, 2 * I, H, generator=g) * 0.05, "down_weight": torch.randn(E, H, I, generator=g) * 0.05,
      "gate_up_channel_scales": torch.ones(H), "down_channel_scales": torch.ones(I),
      "gate_up_quantizer.scale": torch.rand(E * 2 * I * H // G, generator=g) * 0.004 + 0.002,
      "gate_up_quantizer.zero_point_float": -(torch.rand(E * 2 * I * H // G, generator=g) * 18 - 1.5),
      "down_quantizer.scale": torch.rand(E * H * I // G, generator=g) * 0.004 + 0.002,
      "down_quantizer.zero_point_float": -(torch.rand(E * H * I // G, generator=g) * 18 - 1.5)}
for r in range(R):
    st[f"gate_up_pairs_grouped.{r}"] = torch.cat([torch.randperm(G, generator=g) for _ in range(H // G)]).to(torch.int16)
    st[f"gate_up_angles_grouped.{r}"] = torch.zeros(H // 2)
    st[f"down_pairs_grouped.{r}"] = torch.cat([torch.randperm(G, generator=g) for _ in range(I // G)]).to(torch.int16)
    st[f"down_angles_grouped.{r}"] = torch.zeros(I // 2)
proj, rot, bits, group, krot = ref._quantize_moe(st, "cpu")
sd = {"model.layers.3.mlp.experts.gate_up_proj": torch.zeros(1), "model.layers.3.mlp.experts.down_proj": torch.zeros(1), "other": torch.ones(2)}
inj = ref._inject_quantized_moe_state_dict(sd, [(3, "mlp.experts", proj, rot)])
out = {f"in::{k}": v.numpy() for k, v in st.items()}
out.update({f"out::{k}": v.numpy() for k, v in inj.items()})
out["meta"] = np.array([bits, group, krot, E, H, I])
np.savez_compressed(ROOT / "tests" / "golden" / "ref_convert_moe.npz", **out)
print("written", len(inj), "tensors; keys e.g.", sorted(inj)[:4])
