"""MoE expert block on the fused kernels (paroquant_b200/experts.py) vs the oracle: every expert is rotate -> gate | up ->
SiLU * up -> rotate -> down with the block's shared rotations (reference semantics: mlx/modules.py:159-212, RotateSwitchGLU;
format: convert.py:281-406).  Decode batches run each selected expert as ONE two-step chain, larger ones as two launches."""
import numpy as np
import pytest
import torch

from paroquant_b200.checkpoint import ParoLayerBuffers, make_synthetic_activations, make_synthetic_layer

pytestmark = pytest.mark.gpu
TOL = 2e-3       # two fused linears and an activation deep (each stage within 1e-3 of the oracle on the same inputs)


def _blocks(E, H, I):
    shared_gu = make_synthetic_layer(H, [I, I], seed=70)
    shared_d = make_synthetic_layer(I, [H], seed=71)
    blocks = []
    for e in range(E):
        gu, d = make_synthetic_layer(H, [I, I], seed=80 + e), make_synthetic_layer(I, [H], seed=90 + e)
        gu = ParoLayerBuffers(gu.qweight, gu.qzeros, gu.scales, shared_gu.theta, shared_gu.pairs, shared_gu.channel_scales, [I, I], 128, None)
        gu.theta[1], gu.pairs[1], gu.channel_scales[1] = gu.theta[0], gu.pairs[0], gu.channel_scales[0]     # ONE rotation for gate and up
        d = ParoLayerBuffers(d.qweight, d.qzeros, d.scales, shared_d.theta, shared_d.pairs, shared_d.channel_scales, [H], 128, None)
        blocks.append({"gate_up": gu, "down": d})
    return blocks


@pytest.mark.parametrize("T,k", [(1, 2), (5, 2), (40, 3)])
def test_experts_vs_oracle(oracle, T, k):
    from paroquant_b200.experts import ParoExperts

    E, H, I = 4, 256, 384
    blocks = _blocks(E, H, I)
    ex = ParoExperts(blocks, torch.bfloat16)
    g = torch.Generator().manual_seed(T)
    x = make_synthetic_activations(T, H, seed=3 + T, dtype=torch.bfloat16)
    ids = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)])
    w = torch.softmax(torch.randn(T, k, generator=g), -1)
    y = ex(x.cuda(), ids.cuda(), w.cuda()).float().cpu().numpy()
    ref = np.zeros((T, H), np.float64)
    for e in range(E):
        rows, slot = (ids == e).nonzero(as_tuple=True)
        if not len(rows):
            continue
        xe = x[rows].float().numpy()
        act = oracle.silu_and_mul(oracle.linear(xe, blocks[e]["gate_up"].numpy_dict(), "bfloat16"), "bfloat16")
        ye = oracle.linear(act, blocks[e]["down"].numpy_dict(), "bfloat16")
        ref[rows.numpy()] += ye.astype(np.float64) * w[rows, slot].numpy()[:, None].astype(np.float64)
    assert oracle.rel_err(y, oracle.round_to(ref.astype(np.float32), "bfloat16")) < TOL
