"""MoE expert blocks on the fused kernels.

The reference quantises a fused expert block with ONE rotation per projection shared by all experts
(`paroquant/cli/convert.py:281-381`: gate_up rotation on the hidden dim, down rotation on the intermediate dim; stored once as
`<base>.{gate_up,down}_weight_{theta,pairs,channel_scales}`) and serves it only through its MLX backend
(`inference/backends/mlx/modules.py:159-212`, `RotateSwitchGLU`: rotate x, per-expert gate / up, activation, rotate, per-expert
down).  Here every expert is two fused linears -- merged gate|up (two partitions carrying the same rotation) and down -- so an
expert costs two launches, or ONE for decode batches (a two-step chain with SiLU*up folded into the down step).  The rotation is
recomputed per selected expert inside the kernel (it is a prologue of a few microseconds); routing stays in torch.
"""
from __future__ import annotations

import torch

from .chain import ChainStep, ParoChain
from .checkpoint import ParoLayerBuffers
from .linear import ParoLinearKernel


class ParoExperts:
    """Experts of one MoE block.  `blocks[e] = {"gate_up": merged ParoLayerBuffers, "down": ParoLayerBuffers}` as
    `checkpoint_io.load_paro_checkpoint(...).experts[base]` returns them."""

    def __init__(self, blocks: list[dict[str, ParoLayerBuffers]], dtype: torch.dtype = torch.bfloat16, device="cuda", check_pairs: bool = False):
        self.dtype = dtype
        self.gate_up = [ParoLinearKernel.from_buffers(b["gate_up"].to(device), dtype, check_pairs=check_pairs) for b in blocks]
        self.down = [ParoLinearKernel.from_buffers(b["down"].to(device), dtype, check_pairs=check_pairs) for b in blocks]
        self.hidden = self.down[0].shape.out_features
        self.inter = self.down[0].shape.in_features
        self._chains: dict[tuple[int, int], tuple] = {}

    def _chain(self, e: int, m: int):
        """Two-step chain of expert e for m rows over its own fixed buffers (built once per (expert, m))."""
        key = (e, m)
        c = self._chains.get(key)
        if c is None:
            dev = self.gate_up[e].packed.device
            x = torch.empty(m, self.hidden, dtype=self.dtype, device=dev)
            y = torch.empty(m, self.hidden, dtype=self.dtype, device=dev)
            c = (ParoChain([ChainStep(self.gate_up[e], x=x), ChainStep(self.down[e], x_op="silu_mul", y=y)], m), x, y)
            self._chains[key] = c
        return c

    def expert_forward(self, e: int, x: torch.Tensor) -> torch.Tensor:
        """down_e(silu(gate_e(x)) * up_e(x)) for the rows x [m, hidden]."""
        m = x.shape[0]
        if m <= 16:
            ch, xin, y = self._chain(e, m)
            xin.copy_(x)
            ch()
            return y
        gu = self.gate_up[e](x)
        return self.down[e](torch.nn.functional.silu(gu[:, : self.inter]) * gu[:, self.inter:])

    def __call__(self, x: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor) -> torch.Tensor:
        """x [T, hidden], topk_ids / topk_weights [T, k] -> sum_k w_k * expert_{id_k}(x) in fp32, rounded once."""
        out = torch.zeros(x.shape[0], self.hidden, dtype=torch.float32, device=x.device)
        for e in torch.unique(topk_ids).tolist():
            rows, slot = (topk_ids == e).nonzero(as_tuple=True)
            y = self.expert_forward(int(e), x[rows].contiguous())
            out.index_add_(0, rows, y.float() * topk_weights[rows, slot].float()[:, None])
        return out.to(x.dtype)
