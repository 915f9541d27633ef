"""CPU baseline for bench.py: the reference's math on the host cores.  TEST INFRASTRUCTURE --
only bench.py's `cpu_baseline` / `--impl reference` legs import this.

The reference has no CPU execution path (its rotate op is CUDA-only, rotation.cu:133-135, and
ParoQuantHfQuantizer.validate_environment raises without CUDA, quantizer.py:78-80), so the
"reference CPU path" is this port of its arithmetic: AWQ unpack
(inference/backends/mlx/load.py:21-24) -> dequant s*(q-z) (load.py:46-54) -> pairwise rotation
(rotation.cuh:91-173 semantics via torch index ops, fp32) -> torch.matmul, on all host threads.
Two variants: dequantise on every call (true weight-only-quant cost) and dequantise once.
"""
from __future__ import annotations

import os
import time

import torch

_INV = (0, 4, 1, 5, 2, 6, 3, 7)


def unpack(packed: torch.Tensor) -> torch.Tensor:
    w = packed.to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(0, 32, 4, dtype=torch.int64)
    return ((w[:, :, None] >> sh) & 0xF)[:, :, list(_INV)].reshape(packed.shape[0], -1).to(torch.float32)


def dequant(layer, group: int = 128) -> torch.Tensor:
    q, z = unpack(layer.qweight), unpack(layer.qzeros)
    return (q - z.repeat_interleave(group, 0)) * layer.scales.float().repeat_interleave(group, 0)


def rotate(x: torch.Tensor, pairs, theta, cscales, group: int = 128) -> torch.Tensor:
    M, K = x.shape
    v = (x.float() * cscales.float().view(1, K)).clone()
    base = (torch.arange(K) // group * group).view(K // 2, 2)[:, 0]
    for r in range(pairs.shape[0]):
        p = pairs[r].view(K // 2, 2).long()
        i, j = p[:, 0] + base, p[:, 1] + base
        c, s = theta[r].float().cos(), theta[r].float().sin()
        vi, vj = v[:, i], v[:, j]
        v[:, i] = c * vi + s * vj
        v[:, j] = c * vj - s * vi
    return v


def linear(layer, x: torch.Tensor, W: torch.Tensor | None = None) -> torch.Tensor:
    W = dequant(layer, layer.group_size) if W is None else W
    outs, n0 = [], 0
    for p, n in enumerate(layer.part_sizes):
        outs.append(rotate(x, layer.pairs[p], layer.theta[p], layer.channel_scales[p], layer.group_size) @ W[:, n0:n0 + n])
        n0 += n
    return torch.cat(outs, -1)


def time_sample(layers, xs, repeats: int = 1, cached: bool = False) -> float:
    """Seconds for one pass over `layers` (list of ParoLayerBuffers on CPU) with inputs xs."""
    torch.set_num_threads(os.cpu_count() or 1)
    Ws = [dequant(l, l.group_size) for l in layers] if cached else [None] * len(layers)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for l, x, W in zip(layers, xs, Ws):
            linear(l, x, W)
        best = min(best, time.perf_counter() - t0)
    return best
