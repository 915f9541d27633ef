"""Autograd wrapper around `torch.ops.rotation.rotate`.

Surface of /root/reference/paroquant/kernels/cuda/autograd.py:6-65 (`RotateTensorFunc`,
`scaled_pairwise_rotation`) used by the offline optimiser and the checkpoint converter.
Forward is the sm_100a kernel.  Backward walks the rotations in reverse using the identities

    y_i =  c a + s b,   y_j = -s a + c b
    dy_i/dtheta = y_j,  dy_j/dtheta = -y_i          =>  dL/dtheta = sum_rows (G_i y_j - G_j y_i)

on the stage OUTPUT (t) and its gradient (g), then un-rotates both with -theta (a Givens
rotation is orthogonal, so the same kernel back-propagates g).
"""
from __future__ import annotations

import torch


class RotateTensorFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx_ij, theta, scale=None, group_size=128):
        y = torch.ops.rotation.rotate(x, idx_ij, theta, scale, group_size)
        ctx.group_size = group_size
        ctx.has_scale = scale is not None
        ctx.save_for_backward(x, idx_ij, theta, y, *((scale,) if scale is not None else ()))
        return y

    @staticmethod
    def backward(ctx, grad_out):
        x, idx_ij, theta, y = ctx.saved_tensors[:4]
        scale = ctx.saved_tensors[4] if ctx.has_scale else None
        G = ctx.group_size
        krot, K = idx_ij.shape
        rows = y.numel() // K
        t = y.reshape(rows, K)
        g = grad_out.reshape(rows, K).contiguous()
        base = (torch.arange(K, device=idx_ij.device) // G * G).view(K // 2, 2)[:, 0]
        grad_theta = torch.zeros_like(theta)
        for r in reversed(range(krot)):
            pr = idx_ij[r].view(K // 2, 2).long()
            ci, cj = pr[:, 0] + base, pr[:, 1] + base
            grad_theta[r] = ((g[:, ci] * t[:, cj] - g[:, cj] * t[:, ci]).sum(0)).to(theta.dtype)
            inv = -theta[r : r + 1]
            t = torch.ops.rotation.rotate(t, idx_ij[r : r + 1], inv, None, G)
            g = torch.ops.rotation.rotate(g, idx_ij[r : r + 1], inv, None, G)
        if scale is None:
            return g.view_as(x).to(x.dtype), None, grad_theta, None, None
        flat_scale = scale.reshape(-1)
        grad_x = (g * flat_scale.unsqueeze(0)).view_as(x).to(x.dtype)
        grad_scale = (x.reshape(rows, K) * g).sum(0).to(scale.dtype).view_as(scale)
        return grad_x, None, grad_theta, grad_scale, None


def scaled_pairwise_rotation(x, idx_ij, theta, scales=None, group_size=128):
    return RotateTensorFunc.apply(x, idx_ij, theta, scales, group_size)
