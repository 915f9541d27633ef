"""Checkpoint I/O: export arithmetic against golden vectors produced by the reference's
own converter functions (tools/gen_convert_golden.py), on-disk round trip, quantised-module detection, merged projections."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from paroquant_b200 import checkpoint_io as cio
from paroquant_b200.checkpoint import ParoLayerBuffers, make_synthetic_layer, pack_awq, unpack_awq

GOLD = Path(__file__).parent / "golden" / "ref_convert.npz"


def test_export_arithmetic_matches_reference_converter():
    g = np.load(GOLD)
    q, s2, z2 = cio.quantize_rotated(torch.from_numpy(g["weight"]), torch.from_numpy(g["scales_flat"]), torch.from_numpy(g["zp_flat"]),
                                     bits=4, group_size=128)
    assert np.array_equal(q.numpy(), g["quantized"])                   # integer work: bit-exact
    assert np.array_equal(z2.numpy(), g["zeros_2d"])
    assert np.array_equal(s2.numpy(), g["scales_2d"])
    bufs = cio.to_awq_buffers(q, s2, z2)
    for k in ("qweight", "qzeros", "scales"):
        assert np.array_equal(bufs[k].numpy(), g[k]), k
    assert np.array_equal(pack_awq(torch.from_numpy(g["pack_in"])).numpy(), g["pack_out"])
    assert np.array_equal(unpack_awq(torch.from_numpy(g["pack_out"])).numpy(), g["pack_in"])


def _names(prefix, L: ParoLayerBuffers):
    t = {f"{prefix}.qweight": L.qweight, f"{prefix}.qzeros": L.qzeros, f"{prefix}.scales": L.scales, f"{prefix}.theta": L.theta[0],
         f"{prefix}.pairs": L.pairs[0], f"{prefix}.channel_scales": L.channel_scales[0]}
    if L.bias is not None:
        t[f"{prefix}.bias"] = L.bias
    return t


def _write_model(tmp_path):
    q = make_synthetic_layer(256, [256], seed=1, bias=True)
    k = make_synthetic_layer(256, [128], seed=2, bias=True)
    v = make_synthetic_layer(256, [128], seed=3, bias=True)
    o = make_synthetic_layer(256, [256], seed=4)
    tensors = {}
    for name, L in (("q_proj", q), ("k_proj", k), ("v_proj", v), ("o_proj", o)):
        tensors.update(_names(f"model.layers.0.self_attn.{name}", L))
    tensors["model.embed_tokens.weight"] = torch.randn(32, 256).half()
    tensors["model.visual.proj.weight"] = torch.randn(16, 256).half()          # an unquantised linear: left alone
    cio.save_paro_checkpoint(tmp_path, tensors, bits=4, group_size=128, krot=8, base_config={"model_type": "llama"})
    return {"q": q, "k": k, "v": v, "o": o}


def test_round_trip_detection_and_merge(tmp_path):
    src = _write_model(tmp_path)
    cfg = json.loads((tmp_path / "config.json").read_text())
    assert cfg["quantization_config"] == {"quant_method": "paroquant", "bits": 4, "group_size": 128, "krot": 8} and cfg["model_type"] == "llama"
    assert cio.find_quantized_modules(tmp_path) == {f"model.layers.0.self_attn.{n}" for n in ("q_proj", "k_proj", "v_proj", "o_proj")}
    ck = cio.load_paro_checkpoint(tmp_path)
    assert set(ck.dense) == {"model.embed_tokens.weight", "model.visual.proj.weight"}
    got = ck.layers["model.layers.0.self_attn.k_proj"]
    for name in ("qweight", "qzeros", "scales", "theta", "pairs", "channel_scales", "bias"):
        assert torch.equal(getattr(got, name), getattr(src["k"], name)), name
    assert got.part_sizes == [128] and got.krot == 8 and got.in_features == 256

    merged = cio.merged_view(ck)
    assert set(merged) == {"model.layers.0.self_attn.qkv_proj", "model.layers.0.self_attn.o_proj"}
    qkv = merged["model.layers.0.self_attn.qkv_proj"]
    assert qkv.part_sizes == [256, 128, 128] and qkv.out_features == 512 and qkv.theta.shape == (3, 8, 128)
    assert torch.equal(unpack_awq(qkv.qweight)[:, 256:384], unpack_awq(src["k"].qweight))   # partitions keep their columns
    assert torch.equal(qkv.pairs[2], src["v"].pairs[0]) and torch.equal(qkv.bias[384:], src["v"].bias)
    # the merged layer is exactly what the synthetic generator builds for a fused projection of the same parts
    assert qkv.numpy_dict()["part_sizes"] == [256, 128, 128]

    ck2 = cio.load_paro_checkpoint(tmp_path, modules_to_not_convert=["model.layers.0.self_attn.o_proj"])
    assert "model.layers.0.self_attn.o_proj" not in ck2.layers and "model.layers.0.self_attn.o_proj.qweight" in ck2.dense


def test_round_trip_group64(tmp_path):
    """A group_size = 64 checkpoint (convert.py rotates and quantises in groups of `group_size`) loads into group-64 buffers
    that the fused kernels accept (shape check through the C-ABI, no GPU)."""
    from paroquant_b200 import _cabi
    L = make_synthetic_layer(256, [128], group_size=64, seed=21)
    cio.save_paro_checkpoint(tmp_path, _names("model.layers.0.mlp.down_proj", L), bits=4, group_size=64, krot=8)
    ck = cio.load_paro_checkpoint(tmp_path)
    got = ck.layers["model.layers.0.mlp.down_proj"]
    assert got.group_size == 64 and got.qzeros.shape == (4, 16) and got.scales.shape == (4, 128)
    for name in ("qweight", "qzeros", "scales", "theta", "pairs", "channel_scales"):
        assert torch.equal(getattr(got, name), getattr(L, name)), name
    shape = _cabi.make_shape(got.in_features, got.part_sizes, got.group_size, got.krot, torch.float16)
    assert _cabi.packed_bytes(shape) > 2 * 8960                      # 1 block x 2 record groups, two scale / zero sets each
    (tmp_path / "config.json").write_text(json.dumps({"quantization_config": {"quant_method": "paroquant", "bits": 4, "group_size": 32, "krot": 8}}))
    with pytest.raises(ValueError, match="group_size 64 or 128"):
        cio.load_paro_checkpoint(tmp_path)


def test_loader_errors(tmp_path):
    _write_model(tmp_path)
    (tmp_path / "config.json").write_text(json.dumps({"quantization_config": {"quant_method": "awq"}}))
    with pytest.raises(ValueError, match="expected 'paroquant'"):
        cio.load_paro_checkpoint(tmp_path)
    (tmp_path / "config.json").write_text(json.dumps({"quantization_config": {"quant_method": "paroquant", "bits": 3, "group_size": 128, "krot": 8}}))
    with pytest.raises(ValueError, match="INT4 with group_size 64 or 128 only"):
        cio.load_paro_checkpoint(tmp_path)
    (tmp_path / "config.json").write_text(json.dumps({"quantization_config": {"quant_method": "paroquant", "bits": 4, "group_size": 128, "krot": 4}}))
    with pytest.raises(ValueError, match="buffer shapes do not match"):
        cio.load_paro_checkpoint(tmp_path)
    a, b = make_synthetic_layer(256, [128], seed=5, bias=True), make_synthetic_layer(256, [128], seed=6)
    with pytest.raises(ValueError, match="bias"):
        cio.merge_layers([a, b])
    with pytest.raises(ValueError, match="disagree"):
        cio.merge_layers([a, make_synthetic_layer(384, [128], seed=7, bias=True)])


def test_export_layer_is_cuda_only():
    with pytest.raises(RuntimeError, match="CUDA-only"):
        cio.export_layer({"weight": torch.zeros(8, 128)}, device="cpu")


def test_hf_quantizer_glue_replaces_and_loads(tmp_path):
    """transformers loader glue (reference quantizer.py:30-115) on a tiny random Llama: only modules with `.qweight`
    in the checkpoint become RotateQuantizedLinear; the checkpoint then loads into the model strictly."""
    transformers = pytest.importorskip("transformers")
    from safetensors.torch import load_file
    from paroquant_b200.inference.backends.transformers import RotateQuantizedLinear
    from paroquant_b200.inference.backends.transformers import quantizer as hfq

    cfg = transformers.LlamaConfig(hidden_size=256, intermediate_size=384, num_hidden_layers=1, num_attention_heads=4,
                                   num_key_value_heads=2, vocab_size=64, tie_word_embeddings=False)
    model = transformers.LlamaForCausalLM(cfg)
    tensors = {k: v.detach().clone() for k, v in model.state_dict().items()}
    quantised = {}
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Linear) and ".layers." in name and not name.endswith("down_proj"):   # down_proj stays dense
            L = make_synthetic_layer(mod.in_features, [mod.out_features], seed=len(quantised) + 11)
            tensors.pop(f"{name}.weight")
            tensors.update(_names(name, L))
            quantised[name] = L
    cio.save_paro_checkpoint(tmp_path, tensors, krot=8, base_config=cfg.to_dict())

    assert "paroquant" in transformers.quantizers.auto.AUTO_QUANTIZER_MAPPING
    qz = hfq.ParoQuantHfQuantizer(hfq.ParoQuantConfig(bits=4, group_size=128, krot=8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="requires CUDA"):
            qz.validate_environment()
    assert qz.update_dtype(torch.float32) == torch.float16 and qz.update_dtype(torch.bfloat16) == torch.bfloat16
    model.config._name_or_path = str(tmp_path)
    qz._process_model_before_weight_loading(model)
    swapped = {n for n, m in model.named_modules() if isinstance(m, RotateQuantizedLinear)}
    assert swapped == set(quantised)
    assert isinstance(model.get_submodule("model.layers.0.mlp.down_proj"), torch.nn.Linear)
    assert isinstance(model.lm_head, torch.nn.Linear)
    missing, unexpected = model.load_state_dict(load_file(str(tmp_path / "model.safetensors")), strict=True)
    assert not missing and not unexpected
    name = "model.layers.0.self_attn.k_proj"
    got = model.get_submodule(name)
    assert torch.equal(got.qweight, quantised[name].qweight) and torch.equal(got.pairs, quantised[name].pairs[0])
    assert torch.equal(got.channel_scales, quantised[name].channel_scales[0])


# ------------------------------------------------------------------------------------------ MoE expert blocks (convert.py:281-406)
MOE_GOLD = Path(__file__).parent / "golden" / "ref_convert_moe.npz"


def test_moe_export_arithmetic_and_names_match_reference_converter():
    """`moe_expert_buffers` + `moe_state_entries` against the reference's `_quantize_moe` / `_inject_quantized_moe_state_dict`
    (tools/gen_convert_moe_golden.py; rotation stubbed by the identity on both sides: the integer work and the naming)."""
    g = np.load(MOE_GOLD)
    bits, group, krot, E, H, I = (int(v) for v in g["meta"])
    st = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in::")}
    quant = {}
    for tag, w in (("gate_up", st["gate_up_weight"]), ("down", st["down_weight"])):
        quant[tag] = cio.quantize_rotated(w.reshape(-1, w.shape[-1]) * st[f"{tag}_channel_scales"].reshape(1, -1), st[f"{tag}_quantizer.scale"],
                                          st[f"{tag}_quantizer.zero_point_float"], bits=bits, group_size=group)
    proj = cio.moe_expert_buffers(quant["gate_up"], quant["down"], E)
    rot = {}
    for tag in ("gate_up", "down"):
        rot[f"{tag}_weight_theta"] = torch.stack([st[f"{tag}_angles_grouped.{r}"] for r in range(krot)]).to(torch.float16)
        rot[f"{tag}_weight_pairs"] = torch.stack([st[f"{tag}_pairs_grouped.{r}"] for r in range(krot)])
        rot[f"{tag}_weight_channel_scales"] = (1.0 / st[f"{tag}_channel_scales"].reshape(1, -1)).to(torch.float16)
    ours = cio.moe_state_entries("model.layers.3.mlp.experts", proj, rot)
    ref = {k[5:]: g[k] for k in g.files if k.startswith("out::")}
    assert set(ours) == set(ref) - {"other"}                      # the fused gate_up_proj / down_proj tensors are gone, the rest untouched
    for k, v in ours.items():
        assert np.array_equal(v.numpy(), ref[k]), k


def test_moe_round_trip_through_the_loader(tmp_path):
    """Expert modules carry no rotation of their own: the loader attaches the block's shared rotation and merges gate | up."""
    E, H, I = 3, 256, 128
    shared_gu = make_synthetic_layer(H, [I, I], seed=70)          # rotation of the hidden dim, shared by every expert
    shared_d = make_synthetic_layer(I, [H], seed=71)
    tensors, experts = {}, []
    for e in range(E):
        gu = make_synthetic_layer(H, [I, I], seed=80 + e)
        d = make_synthetic_layer(I, [H], seed=90 + e)
        experts.append((gu, d))
        qw, qz = gu.qweight, gu.qzeros
        for proj, sl, L in (("gate_proj", slice(0, I // 8), gu), ("up_proj", slice(I // 8, 2 * I // 8), gu)):
            tensors[f"model.layers.0.mlp.experts.{e}.{proj}.qweight"] = qw[:, sl].contiguous()
            tensors[f"model.layers.0.mlp.experts.{e}.{proj}.qzeros"] = qz[:, sl].contiguous()
            tensors[f"model.layers.0.mlp.experts.{e}.{proj}.scales"] = gu.scales[:, sl.start * 8:sl.stop * 8].contiguous()
        for leaf in ("qweight", "qzeros", "scales"):
            tensors[f"model.layers.0.mlp.experts.{e}.down_proj.{leaf}"] = getattr(d, leaf)
    for tag, L in (("gate_up", shared_gu), ("down", shared_d)):
        tensors[f"model.layers.0.mlp.experts.{tag}_weight_theta"] = L.theta[0]
        tensors[f"model.layers.0.mlp.experts.{tag}_weight_pairs"] = L.pairs[0]
        tensors[f"model.layers.0.mlp.experts.{tag}_weight_channel_scales"] = L.channel_scales[0]
    tensors["model.layers.0.mlp.gate.weight"] = torch.randn(E, H).half()       # the router: dense, untouched
    cio.save_paro_checkpoint(tmp_path, tensors, base_config={"model_type": "qwen3_moe"})
    ck = cio.load_paro_checkpoint(tmp_path)
    assert not ck.layers and set(ck.experts) == {"model.layers.0.mlp.experts"} and set(ck.dense) == {"model.layers.0.mlp.gate.weight"}
    blocks = ck.experts["model.layers.0.mlp.experts"]
    assert len(blocks) == E
    for e, b in enumerate(blocks):
        gu, d = experts[e]
        assert b["gate_up"].part_sizes == [I, I] and b["down"].part_sizes == [H]
        assert torch.equal(b["gate_up"].qweight, gu.qweight) and torch.equal(b["gate_up"].scales, gu.scales)
        assert torch.equal(b["down"].qweight, d.qweight)
        for p in range(2):                                        # both partitions carry the block's shared rotation
            assert torch.equal(b["gate_up"].theta[p], shared_gu.theta[0]) and torch.equal(b["gate_up"].pairs[p], shared_gu.pairs[0])
        assert torch.equal(b["down"].pairs[0], shared_d.pairs[0]) and torch.equal(b["down"].channel_scales.reshape(-1), shared_d.channel_scales[0].reshape(-1))
