"""Tensor-parallel sharding of ParoQuant linears (one process per GPU, torch.distributed).

What the reference does at load time (/root/reference/paroquant/inference/backends/vllm/
plugin.py:33-50,196-198) and what vLLM does around `quant_method.apply`
(vllm/model_executor/layers/linear.py: RowParallelLinear.forward -> tensor_model_parallel_all_reduce):

  column-parallel (q/k/v, gate/up): every partition's N is split across ranks; each rank keeps the
      FULL rotation metadata and rotates the full x redundantly; no communication.
  row-parallel (o_proj, down_proj): K is split in multiples of 128, so rotation groups and
      quantisation groups never straddle ranks; theta / pairs / channel_scales are sliced along
      the input dim with the weights; each rank produces a partial [M, N]; ONE all-reduce (sum).
      Bias is added after the reduce (rank-0 convention of vLLM is equivalent).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .checkpoint import ParoLayerBuffers


def shard_rows(b: ParoLayerBuffers, rank: int, world: int) -> ParoLayerBuffers:
    """Row-parallel shard: input channels [rank*K/world, (rank+1)*K/world)."""
    K, G = b.in_features, b.group_size
    if K % (world * max(G, 128)):   # a weight record always covers 128 channels (two groups when group_size is 64)
        raise ValueError(f"in_features={K} cannot be split {world}-way on {max(G, 128)}-channel group boundaries")
    ks = K // world
    k0 = rank * ks
    return ParoLayerBuffers(
        qweight=b.qweight[k0:k0 + ks].contiguous(), qzeros=b.qzeros[k0 // G:(k0 + ks) // G].contiguous(),
        scales=b.scales[k0 // G:(k0 + ks) // G].contiguous(), theta=b.theta[..., k0 // 2:(k0 + ks) // 2].contiguous(),
        pairs=b.pairs[..., k0:k0 + ks].contiguous(), channel_scales=b.channel_scales[..., k0:k0 + ks].contiguous(),
        part_sizes=list(b.part_sizes), group_size=G, bias=b.bias, extras=dict(b.extras))


def shard_columns(b: ParoLayerBuffers, rank: int, world: int) -> ParoLayerBuffers:
    """Column-parallel shard: every partition's output columns split evenly; rotations replicated."""
    cols, n0 = [], 0
    parts = []
    for n in b.part_sizes:
        if n % (world * 16):
            raise ValueError(f"partition of {n} columns cannot be split {world}-way in multiples of 16")
        w = n // world
        cols.append(torch.arange(n0 + rank * w, n0 + (rank + 1) * w, device=b.qweight.device))
        parts.append(w)
        n0 += n
    col = torch.cat(cols)
    pcol = col.view(-1, 8)[:, 0] // 8           # packed int32 columns (8 outputs per word, whole words move)
    return ParoLayerBuffers(
        qweight=b.qweight[:, pcol].contiguous(), qzeros=b.qzeros[:, pcol].contiguous(),
        scales=b.scales[:, col].contiguous(), theta=b.theta, pairs=b.pairs, channel_scales=b.channel_scales,
        part_sizes=parts, group_size=b.group_size, bias=None if b.bias is None else b.bias[col].contiguous(),
        extras=dict(b.extras))


class RowParallelParoLinear:
    """K-sharded fused linear + all-reduce of the [M, N] partials."""

    def __init__(self, full: ParoLayerBuffers, dtype: torch.dtype, device, group=None):
        from .linear import ParoLinearKernel

        self.group = group
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        shard = shard_rows(full, rank, world).to(device)
        self.kernel = ParoLinearKernel.from_buffers(shard, dtype)
        self.bias = None if full.bias is None else full.bias.to(device=device, dtype=dtype)

    def __call__(self, x_shard: torch.Tensor) -> torch.Tensor:
        y = self.kernel(x_shard)
        dist.all_reduce(y, group=self.group)
        if self.bias is not None:
            y = y + self.bias
        return y


class ColumnParallelParoLinear:
    """N-sharded fused linear, no communication (outputs stay sharded, as in vLLM)."""

    def __init__(self, full: ParoLayerBuffers, dtype: torch.dtype, device, group=None):
        from .linear import ParoLinearKernel

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        shard = shard_columns(full, rank, world).to(device)
        self.kernel = ParoLinearKernel.from_buffers(shard, dtype)
        self.bias = None if shard.bias is None else shard.bias.to(device=device, dtype=dtype)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.kernel(x, self.bias)
