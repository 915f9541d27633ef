"""Checkpoint-format helpers for ParoQuant linears (AWQ INT4 buffers + rotation buffers).

On-disk format (normative: /root/reference/paroquant/cli/convert.py:149-203,240-277 and
paroquant/inference/backends/transformers/modules.py:43-55):

  qweight        [K, N/8]   int32   nibble i of word c holds q[k, 8c + ORDER[i]], ORDER = 0,2,4,6,1,3,5,7
  qzeros         [K/G, N/8] int32   same packing of the zero points
  scales         [K/G, N]   fp16
  theta          [R, K/2]   fp16    angle of pair t of group g in rotation r at [r, g*G/2 + t]
  pairs          [R, K]     int16   local indices (0..G-1); (i, j) of pair t at [r, g*G + 2t], [.. + 2t + 1]
  channel_scales [1, K]     fp16    multiplied into the activations before the rotations
  bias           [N]        fp16    optional

Merged projections (QKV, gate_up) keep ONE qweight/qzeros/scales spanning all partitions along N
and P stacked rotation sets (plugin.py:196-198).

This module only builds / validates such buffers (synthetic layers for tests and bench); the
arithmetic lives in csrc/.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)
_AWQ_INVERSE = (0, 4, 1, 5, 2, 6, 3, 7)


def pack_awq(values: torch.Tensor) -> torch.Tensor:
    """[R, C] integers in 0..15 -> int32 [R, C/8] in AWQ nibble order."""
    if values.shape[-1] % 8:
        raise ValueError("last dim must be a multiple of 8")
    v = values.to(torch.int64).reshape(values.shape[0], -1, 8)
    word = torch.zeros(v.shape[:2], dtype=torch.int64, device=values.device)
    for slot, col in enumerate(AWQ_ORDER):
        word |= (v[:, :, col] & 0xF) << (4 * slot)
    # reinterpret the low 32 bits as a signed int32
    word = torch.where(word >= 2**31, word - 2**32, word)
    return word.to(torch.int32)


def unpack_awq(packed: torch.Tensor) -> torch.Tensor:
    """int32 [R, C/8] -> uint8 [R, C]."""
    w = packed.to(torch.int64) & 0xFFFFFFFF
    shifts = torch.arange(0, 32, 4, device=packed.device, dtype=torch.int64)
    nib = ((w[:, :, None] >> shifts) & 0xF).to(torch.uint8)
    return nib[:, :, list(_AWQ_INVERSE)].reshape(packed.shape[0], -1)


@dataclass
class ParoLayerBuffers:
    """One (possibly merged) ParoQuant linear in checkpoint format."""

    qweight: torch.Tensor
    qzeros: torch.Tensor
    scales: torch.Tensor
    theta: torch.Tensor            # [P, R, K/2] fp16
    pairs: torch.Tensor            # [P, R, K]   int16
    channel_scales: torch.Tensor   # [P, 1, K]   fp16
    part_sizes: list[int]
    group_size: int = 128
    bias: torch.Tensor | None = None
    extras: dict = field(default_factory=dict)

    @property
    def in_features(self) -> int:
        return int(self.qweight.shape[0])

    @property
    def out_features(self) -> int:
        return int(self.qweight.shape[1]) * 8

    @property
    def krot(self) -> int:
        return int(self.theta.shape[1])

    def to(self, device) -> "ParoLayerBuffers":
        mv = lambda t: None if t is None else t.to(device)
        return ParoLayerBuffers(mv(self.qweight), mv(self.qzeros), mv(self.scales), mv(self.theta),
                                mv(self.pairs), mv(self.channel_scales), list(self.part_sizes),
                                self.group_size, mv(self.bias), dict(self.extras))

    def numpy_dict(self) -> dict:
        """Plain numpy view for the CPU oracle (fp16 tensors become fp32 values)."""
        f = lambda t: t.detach().cpu().float().numpy()
        return {
            "qweight": self.qweight.cpu().numpy(), "qzeros": self.qzeros.cpu().numpy(),
            "scales": f(self.scales), "theta": f(self.theta), "pairs": self.pairs.cpu().numpy(),
            "channel_scales": f(self.channel_scales), "part_sizes": list(self.part_sizes),
            "group": self.group_size, "bias": None if self.bias is None else f(self.bias),
        }

    def algorithmic_bytes(self, m: int, act_bytes: int = 2) -> int:
        """canonical checkpoint-format bytes one forward must touch."""
        K, N, G, R, P = self.in_features, self.out_features, self.group_size, self.krot, len(self.part_sizes)
        return (K * N // 2 + (K // G) * N * 2 + (K // G) * N // 2
                + P * (R * K * 2 + R * (K // 2) * 2 + K * 2) + m * K * act_bytes + m * N * act_bytes)


def validate_pairs(pairs: torch.Tensor, group_size: int) -> None:
    """Every (rotation, group) slice must be a permutation of 0..G-1 (optim/rotation.py:33-61);
    anything else is a data race in the reference kernel and undefined here too."""
    p = pairs.reshape(-1, group_size).to(torch.int64)
    if p.min() < 0 or p.max() >= group_size:
        raise ValueError("rotation pair index out of range")
    srt = p.sort(dim=-1).values
    if not torch.equal(srt, torch.arange(group_size, device=p.device).expand_as(srt)):
        raise ValueError("rotation pairs of one (rotation, group) are not a permutation of 0..G-1")


def make_synthetic_layer(in_features: int, part_sizes, *, group_size: int = 128, krot: int = 8,
                         seed: int = 1234, device="cpu", bias: bool = False,
                         theta_std: float = 0.3, theta_uniform_pi: bool = False) -> ParoLayerBuffers:
    """Seeded synthetic layer per SURVEY.md section 8(d)."""
    K, G = in_features, group_size
    part_sizes = [int(p) for p in (part_sizes if isinstance(part_sizes, (list, tuple)) else [part_sizes])]
    N, P = sum(part_sizes), len(part_sizes)
    if K % G or N % 8:
        raise ValueError("K must be a multiple of group_size and N of 8")
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if dev.type == "cuda":
        # random words == iid random nibbles; avoids materialising [K, N] int64 for MLP shapes
        qweight = torch.randint(-2**31, 2**31, (K, N // 8), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
        qzeros = torch.randint(-2**31, 2**31, (K // G, N // 8), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
    else:
        qweight = pack_awq(torch.randint(0, 16, (K, N), generator=g, dtype=torch.int32))
        qzeros = pack_awq(torch.randint(0, 16, (K // G, N), generator=g, dtype=torch.int32))
    scales = ((0.5 + torch.rand(K // G, N, generator=g, device=dev)) * (0.02 / 7.5)).to(torch.float16)
    perm = torch.rand(P, krot, K // G, G, generator=g, device=dev).argsort(dim=-1)
    pairs = perm.reshape(P, krot, K).to(torch.int16)
    if theta_uniform_pi:
        theta = ((torch.rand(P, krot, K // 2, generator=g, device=dev) * 2 - 1) * 3.14159).to(torch.float16)
    else:
        theta = (torch.randn(P, krot, K // 2, generator=g, device=dev) * theta_std).to(torch.float16)
    channel_scales = torch.exp(torch.randn(P, 1, K, generator=g, device=dev) * 0.3).to(torch.float16)
    b = (torch.randn(N, generator=g, device=dev) * 0.1).to(torch.float16) if bias else None
    return ParoLayerBuffers(qweight, qzeros, scales, theta, pairs, channel_scales, part_sizes, G, b)


def make_synthetic_activations(m: int, in_features: int, *, seed: int = 4321, device="cpu",
                               dtype=torch.bfloat16) -> torch.Tensor:
    """x ~ N(0,1) with 1 % of the channels scaled x20 (outlier-like), SURVEY.md section 8(d)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randn(m, in_features, generator=g, device=dev)
    n_out = max(1, in_features // 100)
    ch = torch.randperm(in_features, generator=g, device=dev)[:n_out]
    x[:, ch] *= 20.0
    return x.to(dtype)
