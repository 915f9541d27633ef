"""Real-format checkpoint -> kernels, end to end on the GPU (reference `convert.py:240-277,409-464` writes the
format, `quantizer.py:88-115` / `plugin.py:123-151,196-198` load it):

  optimiser state of q / k / v / o  --export_layer (weight rotation on the GPU)-->  AWQ + rotation buffers  --safetensors + config-->
  load_paro_checkpoint --merged_view (qkv: one weight, stacked rotations)--> ParoLinearKernel --> y

checked against the oracle on the loaded buffers (<= 1e-3) AND against the float model the optimiser state describes:
(x * 1/cs) rotated, times the rotated weight, is x @ W^T up to INT4 rounding -- a wrong export or load shows up as O(1) error."""
import json

import numpy as np
import pytest
import torch

from paroquant_b200 import checkpoint_io as cio
from paroquant_b200.checkpoint import make_synthetic_activations

pytestmark = pytest.mark.gpu


def _opt_state(n_out, k_in, seed, group=128, krot=8):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n_out, k_in, generator=g) * 0.05
    pairs = torch.stack([torch.cat([torch.randperm(group, generator=g) for _ in range(k_in // group)]) for _ in range(krot)]).to(torch.int16)
    theta = torch.randn(krot, k_in // 2, generator=g) * 0.3      # the angle scale of make_synthetic_layer
    cs = 0.5 + torch.rand(1, k_in, generator=g)
    st = {"weight": w, "n_bits": torch.tensor(4), "group_size": torch.tensor(group), "channel_scales": cs}
    for r in range(krot):
        st[f"pairs_grouped.{r}"] = pairs[r]
        st[f"angles_grouped.{r}"] = theta[r]
    return st


def test_export_save_load_run(tmp_path, oracle):
    import paroquant_b200.kernels.cuda  # noqa: F401
    from paroquant_b200.linear import ParoLinearKernel

    K, group = 512, 128
    outs = {"q_proj": 256, "k_proj": 128, "v_proj": 128, "o_proj": 512}
    tensors, states = {}, {}
    for i, (name, n) in enumerate(outs.items()):
        st = _opt_state(n, K, seed=10 + i)
        # the optimiser's per-(row, group) scale / zero point: min-max of the ROTATED weight, as its quantiser would learn them
        rot = torch.ops.rotation.rotate((st["weight"] * st["channel_scales"]).cuda(),
                                        torch.stack([st[f"pairs_grouped.{r}"] for r in range(8)]).cuda(),
                                        torch.stack([st[f"angles_grouped.{r}"] for r in range(8)]).cuda(), None, group).cpu()
        rg = rot.view(n, K // group, group)
        lo, hi = rg.amin(-1), rg.amax(-1)
        scale = ((hi - lo) / 15).clamp_min(1e-6)
        st["quantizer.scale"] = scale.reshape(-1)
        st["quantizer.zero_point_float"] = (lo / scale).reshape(-1)
        states[name] = st
        buffers, bits, g, krot = cio.export_layer(st, device="cuda")
        assert (bits, g, krot) == (4, group, 8)
        for key, t in buffers.items():
            tensors[f"model.layers.0.self_attn.{name}.{key}"] = t
    tensors["model.embed_tokens.weight"] = torch.zeros(8, K, dtype=torch.float16)
    cio.save_paro_checkpoint(tmp_path, tensors, base_config={"model_type": "llama"})
    assert json.loads((tmp_path / "config.json").read_text())["quantization_config"]["quant_method"] == "paroquant"

    merged = cio.merged_view(cio.load_paro_checkpoint(tmp_path))
    x = make_synthetic_activations(5, K, seed=3, dtype=torch.bfloat16)
    for mname, parts in (("model.layers.0.self_attn.qkv_proj", ["q_proj", "k_proj", "v_proj"]), ("model.layers.0.self_attn.o_proj", ["o_proj"])):
        L = merged[mname]
        assert L.part_sizes == [outs[p] for p in parts]
        for M in (1, 5):
            y = ParoLinearKernel.from_buffers(L.to("cuda"), torch.bfloat16)(x[:M].cuda()).float().cpu().numpy()
            ref = oracle.linear(x[:M].float().numpy(), L.numpy_dict(), "bfloat16")
            assert oracle.rel_err(y, ref) < 1e-3
            dense = np.concatenate([x[:M].float().numpy() @ states[p]["weight"].numpy().T for p in parts], axis=-1)
            assert oracle.rel_err(y, dense) < 0.25, "export / load changed the function (beyond INT4 rounding)"
