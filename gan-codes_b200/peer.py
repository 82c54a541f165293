"""Data-parallel gradient exchange over NVLink peer memory: host side of csrc/peer_reduce.cu.

One process per GPU.  `PeerGroup.empty(n)` is a COLLECTIVE call: every rank allocates an fp32 buffer of
n elements with the library (cudaMalloc, zeroed), the IPC handles travel through `torch.distributed` (plumbing: one
all_gather_object at set-up, nothing per iteration) and every rank maps every peer's buffer.  `all_reduce(view)` then
launches ONE kernel on the current stream that sums the range across the ranks in place -- no NCCL call, no host
synchronisation, capturable in the iteration's CUDA graph (engine.py).  The sum is formed by one rank per slice in
rank order, so every rank ends with bit-identical values (what N reference replicas with summed gradients hold;
the reference's optimiser steps: deep_fusion_gan/model.py:195-197, 208-210, 226-228).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

from . import _lib

PEER_CTAS = int(os.environ.get("DFGAN_PEER_CTAS", "32"))
PEER_TIMEOUT_MS = int(os.environ.get("DFGAN_PEER_TIMEOUT_MS", "20000"))


class _Raw:
    """__cuda_array_interface__ carrier: lets torch view library-owned device memory without copying."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}


class _Alloc:
    def __init__(self, ptr: int, nbytes: int, peers: List[int]):
        self.ptr, self.nbytes, self.peers = ptr, nbytes, peers


class PeerGroup:
    def __init__(self, device: torch.device, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("PeerGroup needs an initialised torch.distributed process group")
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not 2 <= self.world <= 8:
            raise RuntimeError(f"PeerGroup: 2..8 ranks of one node, got {self.world}")
        self.device = torch.device(device)
        self.lib = _lib.lib()
        self._allocs: List[_Alloc] = []
        self._keep: list = []
        with torch.cuda.device(self.device):
            self._flags = self._alloc_shared(int(self.lib.dfgan_peer_flag_bytes()))
        self._flag_arr = (C.c_void_p * self.world)(*self._flags.peers)

    # ------------------------------------------------------------------------------------ set-up (collective)
    def _alloc_shared(self, nbytes: int) -> _Alloc:
        ptr, mine, err = C.c_void_p(), (None, nbytes), None
        hb = int(self.lib.dfgan_peer_handle_bytes())
        try:
            _lib.check(self.lib.dfgan_peer_alloc(C.c_size_t(nbytes), C.byref(ptr)), "dfgan_peer_alloc")
            handle = (C.c_ubyte * hb)()
            _lib.check(self.lib.dfgan_peer_export(ptr, handle), "dfgan_peer_export")
            mine = (bytes(handle), nbytes)
        except Exception as e:  # noqa: BLE001  (every rank still takes part in the collectives below)
            err = e
        everyone: list = [None] * self.world
        self.dist.all_gather_object(everyone, mine, group=self.group)
        peers = []
        try:
            if err is not None:
                raise err
            for r, (h, nb) in enumerate(everyone):
                if h is None:
                    raise RuntimeError(f"PeerGroup: rank {r} could not allocate / export its buffer")
                if nb != nbytes:
                    raise RuntimeError(f"PeerGroup: rank {r} allocates {nb} bytes, this rank {nbytes}")
                if r == self.rank:
                    peers.append(ptr.value)
                    continue
                q = C.c_void_p()
                buf = (C.c_ubyte * hb).from_buffer_copy(h)
                _lib.check(self.lib.dfgan_peer_import(buf, C.byref(q)), f"dfgan_peer_import (rank {r})")
                peers.append(q.value)
        except Exception as e:  # noqa: BLE001  (reported below, on every rank at once)
            err = e
        # one collective decides for everybody (and is the barrier: nobody proceeds -- and possibly signals into a
        # flag block -- before every rank has mapped everything); a rank that failed makes ALL ranks raise
        okay = torch.tensor([0 if err is not None else 1], device=self.device, dtype=torch.int32)
        self.dist.all_reduce(okay, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(okay.item()) != 1:
            raise RuntimeError(f"PeerGroup: mapping the peers' buffers failed ({err or 'on another rank'})")
        a = _Alloc(ptr.value, nbytes, peers)
        self._allocs.append(a)
        return a

    def empty(self, numel: int) -> torch.Tensor:
        """fp32 buffer of `numel` elements (zeroed) that every rank of the group can reach.  Collective."""
        with torch.cuda.device(self.device):
            nbytes = ((numel * 4 + 15) // 16) * 16
            a = self._alloc_shared(nbytes)
            raw = _Raw(a.ptr, numel)
            t = torch.as_tensor(raw, device=self.device)
        if t.data_ptr() != a.ptr or t.dtype != torch.float32:
            raise RuntimeError("PeerGroup.empty: torch copied the library buffer instead of viewing it")
        self._keep.append(raw)
        return t

    # ------------------------------------------------------------------------------------ per iteration
    def all_reduce(self, t: torch.Tensor, channel: int = 0, ctas: Optional[int] = None, threads: int = 0) -> None:
        """In-place sum of `t` (a contiguous fp32 view into a buffer from `empty`, 16-byte aligned, numel % 4 == 0)
        across the ranks, on the current stream.  Every rank issues the same calls per channel in the same order.
        threads <= 128 selects the small-footprint kernel shape for an exchange that overlaps other kernels."""
        p, n = t.data_ptr(), t.numel()
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("PeerGroup.all_reduce: contiguous fp32 tensor expected")
        for a in self._allocs:
            if a.ptr <= p and p + 4 * n <= a.ptr + a.nbytes:
                off = p - a.ptr
                bufs = (C.c_void_p * self.world)(*[q + off for q in a.peers])
                st = torch.cuda.current_stream(self.device).cuda_stream
                _lib.check(self.lib.dfgan_peer_allreduce_f32(bufs, self._flag_arr, self.world, self.rank, n, channel,
                                                             ctas or PEER_CTAS, threads, PEER_TIMEOUT_MS,
                                                             C.c_void_p(st)),
                           "dfgan_peer_allreduce_f32")
                return
        raise RuntimeError("PeerGroup.all_reduce: tensor is not inside a buffer of this group")

    def close(self) -> None:
        """Unmap the peers' buffers and free the local ones (every rank, after a device synchronisation)."""
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)
        for a in self._allocs:
            for r, q in enumerate(a.peers):
                if r != self.rank:
                    self.lib.dfgan_peer_close(C.c_void_p(q))
        self.dist.barrier(group=self.group)
        for a in self._allocs:
            self.lib.dfgan_peer_free(C.c_void_p(a.ptr))
        self._allocs = []
