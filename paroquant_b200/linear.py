"""Fused ParoQuant linear: the object both module surfaces (HF `RotateQuantizedLinear`, vLLM
`ParoQuantLinearMethod`) delegate to.  One C-ABI call per forward
(include/paro_b200.h: paro_linear_forward), exposed to torch.compile / CUDA graphs as

    paro::linear(Tensor x, Tensor packed, Tensor(a!) workspace, Tensor? bias, int[] meta) -> Tensor

meta = [in_features, group_size, krot, dtype_code, *part_sizes].
"""
from __future__ import annotations

import torch

from . import _cabi
from .checkpoint import ParoLayerBuffers, validate_pairs

torch.library.define(
    "paro::linear", "(Tensor x, Tensor packed, Tensor(a!) workspace, Tensor? bias, int[] meta) -> Tensor")

_shape_cache: dict[tuple, _cabi.ParoLinearShape] = {}


def _shape_from_meta(meta) -> _cabi.ParoLinearShape:
    key = tuple(int(v) for v in meta)
    s = _shape_cache.get(key)
    if s is None:
        s = _cabi.make_shape(key[0], key[4:], key[1], key[2], _cabi._CODE_DTYPE[key[3]])
        _shape_cache[key] = s
    return s


@torch.library.impl("paro::linear", "CUDA")
def _linear_cuda(x, packed, workspace, bias, meta):
    return _cabi.linear_forward(_shape_from_meta(meta), packed, x, bias, workspace)


@torch.library.register_fake("paro::linear")
def _linear_fake(x, packed, workspace, bias, meta):
    return x.new_empty(*x.shape[:-1], sum(int(v) for v in meta[4:]))


# paro_linear_forward treats its workspace as pure scratch (the rotated activations of the M > 16 path), and the linears
# of a model run one after another on a stream: ONE buffer per device serves them all (a private one per layer would be
# n_parts * M * K * 2 bytes each -- gigabytes for a 32-layer model at 4096 tokens).  Buffers are only ever replaced by
# larger ones and the old ones stay alive, so CUDA graphs captured earlier keep valid pointers.
_shared_scratch: dict[torch.device, list[torch.Tensor]] = {}


def shared_workspace(device, nbytes: int) -> torch.Tensor:
    dev = torch.device(device)
    bufs = _shared_scratch.setdefault(dev, [])
    if not bufs or bufs[-1].numel() < nbytes:
        bufs.append(torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=dev))
    return bufs[-1]


class ParoLinearKernel:
    """Prepacked weights + rotation metadata of one (possibly merged) linear on one GPU.
    `private_workspace=True` gives the layer its own scratch (needed only when two linears may run concurrently on
    different streams; include/paro_b200.h: a workspace must not be shared by concurrent launches)."""

    def __init__(self, packed: torch.Tensor, shape: _cabi.ParoLinearShape, max_m: int = 16, private_workspace: bool = False):
        self.packed = packed
        self.shape = shape
        self.meta = [shape.in_features, shape.group_size, shape.krot, shape.dtype,
                     *list(shape.part_sizes[: shape.n_parts])]
        self.private_workspace = private_workspace
        self.max_m = 0
        self.workspace = None
        self._ensure_workspace(max_m)

    @classmethod
    def from_tensors(cls, qweight, qzeros, scales, theta, pairs, channel_scales, part_sizes, *,
                     group_size: int = 128, dtype: torch.dtype = torch.bfloat16, check_pairs: bool = True,
                     max_m: int = 16, private_workspace: bool = False) -> "ParoLinearKernel":
        """theta [P,R,K/2], pairs [P,R,K], channel_scales [P,1,K] or [P,K] (2-D inputs mean P = 1)."""
        if theta.dim() == 2:
            theta, pairs, channel_scales = theta[None], pairs[None], channel_scales.reshape(1, -1)
        _cabi._need_cuda(qweight, qzeros, scales, theta, pairs, channel_scales)
        if check_pairs:
            validate_pairs(pairs, group_size)
        shape = _cabi.make_shape(qweight.shape[0], part_sizes, group_size, theta.shape[1], dtype)
        packed = _cabi.prepack(shape, qweight, qzeros, scales, pairs, theta, channel_scales)
        return cls(packed, shape, max_m, private_workspace)

    @classmethod
    def from_buffers(cls, b: ParoLayerBuffers, dtype: torch.dtype = torch.bfloat16, **kw) -> "ParoLinearKernel":
        return cls.from_tensors(b.qweight, b.qzeros, b.scales, b.theta, b.pairs, b.channel_scales, b.part_sizes,
                                group_size=b.group_size, dtype=dtype, **kw)

    def _ensure_workspace(self, m: int) -> None:
        if self.workspace is None or m > self.max_m:
            if self.private_workspace:
                self.workspace = _cabi.new_workspace(self.shape, m, self.packed.device)
            else:
                self.workspace = shared_workspace(self.packed.device, _cabi.workspace_bytes(self.shape, m))
            self.max_m = m

    def __call__(self, x: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
        m = x.numel() // self.shape.in_features
        if m > self.max_m:
            self._ensure_workspace(m)
        return torch.ops.paro.linear(x, self.packed, self.workspace, bias, self.meta)

    def dense_weight(self) -> torch.Tensor:
        """[K, N] dequantised operand T((q - z) * s) exactly as the kernels form it (tests)."""
        return _cabi.unpack_dense(self.shape, self.packed)
