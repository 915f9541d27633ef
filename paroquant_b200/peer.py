"""Peer-mapped device buffers for the in-kernel tensor-parallel sum (one process per GPU, one node, NVLink).

A row-parallel step of a chain (include/paro_b200.h: paro_tp_info) writes its block sums into a buffer on EVERY rank; each
rank therefore needs device addresses, valid in its own process, of all ranks' buffers.  torch.distributed is plumbing here:
`torch.distributed._symmetric_memory` when it works, else CUDA IPC handles of a plain allocation exchanged with
all_gather_object.  The buffer is zero-filled (tags of the kernels start at 1).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class PeerBuffer:
    """`local` is this rank's buffer (uint8); `ptrs[r]` the address of rank r's buffer as seen from this process."""

    def __init__(self, nbytes: int, device, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        self._keep = []
        nbytes = (int(nbytes) + 255) // 256 * 256
        try:
            self._symmetric(nbytes)
            self.how = "symmetric_memory"
        except Exception as e:  # noqa: BLE001 -- any failure of the experimental API falls back to plain CUDA IPC
            self._ipc(nbytes)
            self.how = f"cuda_ipc (symmetric memory unavailable: {type(e).__name__})"
        dist.barrier(group)

    def _symmetric(self, nbytes: int) -> None:
        import torch.distributed._symmetric_memory as symm

        t = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        t.zero_()
        torch.cuda.synchronize(self.device)
        name = (self.group or dist.group.WORLD).group_name
        try:
            hdl = symm.rendezvous(t, group=name)
        except TypeError:
            hdl = symm.rendezvous(t, name)
        self.local, self.ptrs = t, [int(p) for p in hdl.buffer_ptrs]
        self._keep.append(hdl)
        if len(self.ptrs) != self.world or self.ptrs[self.rank] != t.data_ptr():
            raise RuntimeError("unexpected symmetric-memory handle layout")

    def _ipc(self, nbytes: int) -> None:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        torch.cuda.synchronize(self.device)
        st = t.untyped_storage()
        meta = st._share_cuda_()
        off = t.data_ptr() - st.data_ptr()
        metas = [None] * self.world
        dist.all_gather_object(metas, (meta, off), self.group)
        self.local, self.ptrs = t, []
        for r, (m, o) in enumerate(metas):
            if r == self.rank:
                self.ptrs.append(t.data_ptr())
            else:
                peer = torch.UntypedStorage._new_shared_cuda(*m)
                self._keep.append(peer)
                self.ptrs.append(peer.data_ptr() + o)
