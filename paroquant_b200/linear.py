"""Fused ParoQuant linear: the object both module surfaces (HF `RotateQuantizedLinear`, vLLM
`ParoQuantLinearMethod`) delegate to.  One C-ABI call per forward
(include/paro_b200.h: paro_linear_forward), exposed to torch.compile / CUDA graphs as

    paro::linear(Tensor x, Tensor packed, Tensor? workspace, Tensor? bias, int[] meta) -> Tensor

meta = [in_features, group_size, krot, dtype_code, *part_sizes].
"""
from __future__ import annotations

import torch

from . import _cabi
from .checkpoint import ParoLayerBuffers, validate_pairs

torch.library.define(
    "paro::linear", "(Tensor x, Tensor packed, Tensor? workspace, Tensor? bias, int[] meta) -> Tensor")
# `workspace` is scratch (plus the launch epoch, which only the kernels read): it is NOT declared as mutated -- a mutable optional
# argument sends torch.compile through auto_functionalized, which Inductor (2.11) fails to lower for this op; launches are stream
# ordered anyway, so nothing can observe the scratch between two calls

_shape_cache: dict[tuple, _cabi.ParoLinearShape] = {}


def _shape_from_meta(meta) -> _cabi.ParoLinearShape:
    key = tuple(int(v) for v in meta)
    s = _shape_cache.get(key)
    if s is None:
        s = _cabi.make_shape(key[0], key[4:], key[1], key[2], _cabi._CODE_DTYPE[key[3]])
        _shape_cache[key] = s
    return s


_bias_casts: dict[tuple, tuple] = {}


def _bias_as(bias, dtype):
    """Bias in the activation dtype; a checkpoint's fp16 bias meeting bf16 activations is cast ONCE (per bias tensor and version),
    not per forward (the reference pays a cast kernel on every call)."""
    if bias is None or (bias.dtype == dtype and bias.is_contiguous()):
        return bias
    key = (bias.data_ptr(), dtype)
    hit = _bias_casts.get(key)
    if hit is None or hit[0] != bias._version or hit[1].shape != bias.shape:
        if len(_bias_casts) > 4096:
            _bias_casts.clear()
        hit = (bias._version, bias.to(dtype).contiguous())
        _bias_casts[key] = hit
    return hit[1]


@torch.library.impl("paro::linear", "CUDA")
def _linear_cuda(x, packed, workspace, bias, meta):
    # Everything that looks at sizes lives HERE, inside the opaque op: torch.compile / vLLM trace `__call__` with a symbolic
    # batch dimension and must see a branch-free call (no ctypes, no `m > max_m`).
    shape = _shape_from_meta(meta)
    bias = _bias_as(bias, x.dtype)
    if workspace is None:
        m = x.numel() // shape.in_features
        workspace = shared_workspace(x.device, _cabi.workspace_bytes(shape, max(m, 1)))
    return _cabi.linear_forward(shape, packed, x, bias, workspace)


@torch.library.register_fake("paro::linear")
def _linear_fake(x, packed, workspace, bias, meta):
    return x.new_empty(*x.shape[:-1], sum(int(v) for v in meta[4:]))


# paro_linear_forward's workspace is scratch (partial slots of the M <= 16 kernel, rotated activations of the M > 16 path)
# behind a 256-byte head that holds the launch epoch, and the linears of a model run one after another on a stream: ONE
# buffer per device serves them all (a private one per layer would be n_parts * M * K * 2 bytes each -- gigabytes for a
# 32-layer model at 4096 tokens).  It grows geometrically (at least 2x), so a workload whose batch creeps up allocates
# O(log) buffers; superseded buffers stay alive because CUDA graphs captured earlier hold their addresses, and their sizes
# sum to less than the current one.  `reserve()` sizes it once up front (vLLM: max_num_batched_tokens).
_shared_scratch: dict[torch.device, list[torch.Tensor]] = {}


def shared_workspace(device, nbytes: int) -> torch.Tensor:
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    bufs = _shared_scratch.setdefault(dev, [])
    if not bufs or bufs[-1].numel() < nbytes:
        grow = max(int(nbytes), 2 * bufs[-1].numel() if bufs else 0, 1 << 20)
        bufs.append(torch.zeros(grow, dtype=torch.uint8, device=dev))
    return bufs[-1]


def reserve(device, shapes, max_tokens: int) -> None:
    """Size the shared scratch once for every linear shape of a model at `max_tokens` rows."""
    need = max(_cabi.workspace_bytes(s, max_tokens) for s in shapes)
    shared_workspace(device, need)


class ParoLinearKernel:
    """Prepacked weights + rotation metadata of one (possibly merged) linear on one GPU.
    `private_workspace=True` gives the layer its own scratch sized for `max_m` rows (needed only when two linears may run
    concurrently on different streams; include/paro_b200.h: a workspace must not be shared by concurrent launches);
    otherwise the per-device shared scratch is used and sized inside the op."""

    def __init__(self, packed: torch.Tensor, shape: _cabi.ParoLinearShape, max_m: int = 16, private_workspace: bool = False):
        self.packed = packed
        self.shape = shape
        self.meta = [int(shape.in_features), int(shape.group_size), int(shape.krot), int(shape.dtype),
                     *[int(v) for v in shape.part_sizes[: shape.n_parts]]]
        self.dtype = _cabi._CODE_DTYPE[shape.dtype]
        self.private_workspace = private_workspace
        self._private = _cabi.new_workspace(shape, max_m, packed.device) if private_workspace else None
        shared_workspace(packed.device, _cabi.workspace_bytes(shape, max_m))   # warm the shared scratch outside any capture

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

    @property
    def workspace(self) -> torch.Tensor:
        """The scratch the next launch will use (tools and benches that call the C-ABI directly)."""
        return self._private if self._private is not None else shared_workspace(self.packed.device, 0)

    def __call__(self, x: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
        return torch.ops.paro.linear(x, self.packed, self._private, bias, self.meta)   # nothing here looks at sizes or dtypes

    def forward_into(self, x: torch.Tensor, out: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
        """Same as __call__ but writes into a caller-owned [.., N] buffer (fixed addresses for CUDA graphs / chains)."""
        m = x.numel() // self.shape.in_features
        ws = self._private if self._private is not None else shared_workspace(x.device, _cabi.workspace_bytes(self.shape, max(m, 1)))
        return _cabi.linear_forward(self.shape, self.packed, x, _bias_as(bias, self.dtype), ws, out=out)

    def dense_weight(self) -> torch.Tensor:
        """[K, N] dequantised operand T((q - z) * s) exactly as the kernels form it (tests)."""
        return _cabi.unpack_dense(self.shape, self.packed)
