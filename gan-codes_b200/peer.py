"""Data-parallel gradient exchange over NVLink peer memory: host side of csrc/peer_reduce.cu.

One process per GPU.  `PeerGroup.empty(n)` is a COLLECTIVE call: every rank allocates an fp32 buffer of
n elements with the library (cudaMalloc, zeroed), the IPC handles travel through `torch.distributed` (plumbing: one
all_gather_object at set-up, nothing per iteration) and every rank maps every peer's buffer.  `all_reduce(view)` then
launches ONE kernel on the current stream that sums the range across the ranks in place -- no NCCL call, no host
synchronisation, capturable in the iteration's CUDA graph (engine.py).  When the NVSwitch offers multicast (NVLS) the
buffers come from torch's symmetric-memory allocator instead (allocation + multicast binding are plumbing) and the
kernel reduces inside the switch (`dfgan_peer_allreduce_mc_f32`).  The sum is formed by one rank per slice in
rank order, so every rank ends with bit-identical values (what N reference replicas with summed gradients hold;
the reference's optimiser steps: deep_fusion_gan/model.py:195-197, 208-210, 226-228).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import List, Optional

import torch

from . import _lib

PEER_CTAS = int(os.environ.get("DFGAN_PEER_CTAS", "32"))
# bound of a wait on a peer inside an exchange kernel (nominal: the kernels count polls of ~0.5-1 us); generous, because a
# rank may legitimately arrive late (first eager iteration, a checkpoint written by rank 0 between epochs)
PEER_TIMEOUT_MS = int(os.environ.get("DFGAN_PEER_TIMEOUT_MS", "120000"))
# In-switch reduction (NVLS): the gradient buffers come from torch's symmetric-memory allocator (CUDA VMM allocations
# bound to a multicast object) and the exchange is reduced inside the NVSwitch (multimem.ld_reduce / multimem.st): a GPU
# then moves (1 + 1/N) x the range per direction instead of 2 (N-1)/N x.  Measured (tools/peer_check.py, bit-identical on
# all ranks, 50 graph replays): 78 MB in 195 us against 243 us unicast at 8 GPUs, but 214 us against 137 us at 2 GPUs
# (1.5x MORE bytes there); iteration at 4 GPUs 11.20 ms against 11.32 ms, at 8 GPUs 11.20 ms against 11.39-11.55 ms.
# "auto" (default): in-switch from 4 ranks on; "1": at any rank count; "0", or any rank failing to set it up: library
# cudaMalloc + CUDA IPC buffers and the unicast pull / push kernel.
PEER_MULTIMEM = os.environ.get("DFGAN_PEER_MULTIMEM", "auto")


def owned_slice(lo: int, hi: int, rank: int, world: int):
    """[lo', hi') = the part of elements [lo, hi) of a flat fp32 buffer that `rank` of `world` reduces (reduce-scatter),
    updates (sharded Adam) and pushes (all-gather): the partition csrc/peer_reduce.cu uses -- ceil(n4 / world) 16-byte
    elements per rank, n4 = (hi - lo) / 4, the last ranks possibly empty."""
    n4 = (hi - lo) // 4
    per = (n4 + world - 1) // world
    return lo + 4 * min(per * rank, n4), lo + 4 * min(per * (rank + 1), n4)


class _Raw:
    """__cuda_array_interface__ carrier: lets torch view library-owned device memory without copying."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}


class _Alloc:
    def __init__(self, ptr: int, nbytes: int, peers: List[int], mc: int = 0, owner=None):
        # mc: multicast address of the buffer (0 = none); owner: what keeps a torch-allocated buffer alive
        self.ptr, self.nbytes, self.peers, self.mc, self.owner = ptr, nbytes, peers, mc, owner


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
        self.multimem = PEER_MULTIMEM == "1" or (PEER_MULTIMEM == "auto" and self.world >= 4)
        self._slots = None
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

    def _agree(self, ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(t.item()) == 1

    def _empty_symm(self, numel: int) -> Optional[torch.Tensor]:
        """Buffer from torch's symmetric-memory allocator, mapped on every rank and -- with NVLS -- bound to a multicast
        object.  Returns None (on EVERY rank) when any rank cannot provide a multicast mapping."""
        t, a, why = None, None, ""
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else self.dist.group.WORLD
            padded = ((numel + 3) // 4) * 4
            t = symm.empty(padded, dtype=torch.float32, device=self.device)
            hdl = symm.rendezvous(t, group)
            t.zero_()
            peers = [int(q) for q in hdl.buffer_ptrs]
            mc = int(hdl.multicast_ptr or 0)     # 0: the allocator could not bind a multicast object
            if mc == 0:
                raise RuntimeError("no multicast mapping")
            if len(peers) != self.world or peers[self.rank] != t.data_ptr():
                raise RuntimeError("unexpected symmetric-memory pointer table")
            a = _Alloc(t.data_ptr(), padded * 4, peers, mc, owner=(t, hdl))
        except Exception as e:  # noqa: BLE001  (decided collectively below)
            why = f"{type(e).__name__}: {e}"
        if not self._agree(a is not None):
            if self.rank == 0:
                print(f"[dfgan] NVLS multicast buffers unavailable ({why or 'on another rank'}): CUDA IPC + unicast "
                      "exchange", file=sys.stderr, flush=True)
            self.multimem = False
            return None
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)
        self._allocs.append(a)
        return t[:numel]

    def empty(self, numel: int) -> torch.Tensor:
        """fp32 buffer of `numel` elements (zeroed) that every rank of the group can reach.  Collective."""
        if self.multimem:
            t = self._empty_symm(numel)
            if t is not None:
                return t
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
    def _locate(self, t: torch.Tensor, what: str):
        """(allocation, byte offset) of a contiguous fp32 view into a buffer from `empty`."""
        p, n = t.data_ptr(), t.numel()
        if t.dtype != torch.float32 or not t.is_contiguous() or n % 4 or p % 16:
            raise RuntimeError(f"PeerGroup.{what}: contiguous fp32 tensor, 16-byte aligned, numel % 4 == 0 expected")
        for a in self._allocs:
            if a.ptr <= p and p + 4 * n <= a.ptr + a.nbytes:
                return a, p - a.ptr
        raise RuntimeError(f"PeerGroup.{what}: tensor is not inside a buffer of this group")

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def all_reduce(self, t: torch.Tensor, channel: int = 0, ctas: Optional[int] = None, threads: int = 0) -> None:
        """In-place sum of `t` (a contiguous fp32 view into a buffer from `empty`, 16-byte aligned, numel % 4 == 0)
        across the ranks, on the current stream.  Every rank issues the same calls per channel in the same order.
        threads <= 128 selects the small-footprint kernel shape for an exchange that overlaps other kernels."""
        a, off = self._locate(t, "all_reduce")
        n = t.numel()
        if a.mc and self.multimem:
            _lib.check(self.lib.dfgan_peer_allreduce_mc_f32(C.c_void_p(a.mc + off), self._flag_arr, self.world,
                                                            self.rank, n, channel, ctas or PEER_CTAS, threads,
                                                            PEER_TIMEOUT_MS, self._stream()),
                       "dfgan_peer_allreduce_mc_f32")
            return
        bufs = (C.c_void_p * self.world)(*[q + off for q in a.peers])
        _lib.check(self.lib.dfgan_peer_allreduce_f32(bufs, self._flag_arr, self.world, self.rank, n, channel,
                                                     ctas or PEER_CTAS, threads, PEER_TIMEOUT_MS, self._stream()),
                   "dfgan_peer_allreduce_f32")

    # ---- sharded optimiser step: reduce-scatter -> (norm of the owned slices, summed over ranks) -> all-gather
    def owned(self, lo: int, hi: int, rank: Optional[int] = None):
        """[lo', hi') = the slice of elements [lo, hi) that `rank` (default: this rank) reduces / updates / pushes --
        the partition the kernels use: ceil(n4 / world) 16-byte elements per rank."""
        return owned_slice(lo, hi, self.rank if rank is None else rank, self.world)

    def reduce_scatter(self, t: torch.Tensor, channel: int = 0, ctas: Optional[int] = None, threads: int = 0) -> None:
        """Like all_reduce, but only this rank's slice of `t` (see `owned`) receives the sum; the rest of `t` keeps the
        local values."""
        a, off = self._locate(t, "reduce_scatter")
        bufs = (C.c_void_p * self.world)(*[q + off for q in a.peers])
        _lib.check(self.lib.dfgan_peer_reduce_scatter_f32(bufs, self._flag_arr, self.world, self.rank, t.numel(), channel,
                                                          ctas or PEER_CTAS, threads, PEER_TIMEOUT_MS, self._stream()),
                   "dfgan_peer_reduce_scatter_f32")

    def all_gather(self, t: torch.Tensor, channel: int = 3, ctas: Optional[int] = None) -> None:
        """Every rank pushes its slice of `t` into every peer's copy of `t`; when the kernel has finished, all slices
        of all peers have landed here."""
        a, off = self._locate(t, "all_gather")
        bufs = (C.c_void_p * self.world)(*[q + off for q in a.peers])
        _lib.check(self.lib.dfgan_peer_allgather_f32(bufs, self._flag_arr, self.world, self.rank, t.numel(), channel,
                                                     ctas or PEER_CTAS, PEER_TIMEOUT_MS, self._stream()),
                   "dfgan_peer_allgather_f32")

    def sum_scalar(self, partial: torch.Tensor, total: torch.Tensor, channel: int = 2) -> None:
        """total[0] = sum over the ranks (rank order, bit-identical everywhere) of partial[0]; device scalars."""
        if self._slots is None:
            raise RuntimeError("PeerGroup.sum_scalar: call enable_scalar_exchange() (collective) at set-up")
        _lib.check(self.lib.dfgan_peer_sum_scalar_f32(C.c_void_p(partial.data_ptr()), self._slot_arr,
                                                      C.c_void_p(total.data_ptr()), self._flag_arr, self.world, self.rank,
                                                      channel, PEER_TIMEOUT_MS, self._stream()),
                   "dfgan_peer_sum_scalar_f32")

    def enable_scalar_exchange(self) -> None:
        """Collective: the per-rank slot array `sum_scalar` writes into."""
        if self._slots is None:
            with torch.cuda.device(self.device):
                self._slots = self._alloc_shared(64)
            self._slot_arr = (C.c_void_p * self.world)(*self._slots.peers)

    def peer_view(self, t: torch.Tensor, rank: int) -> torch.Tensor:
        """`t` (a view into a buffer from `empty`) as stored by `rank`: a tensor over the peer-mapped address."""
        a, off = self._locate(t, "peer_view")
        raw = _Raw(a.peers[rank] + off, t.numel())
        v = torch.as_tensor(raw, device=self.device)
        self._keep.append(raw)
        return v

    def close(self) -> None:
        """Unmap the peers' buffers and free the local ones (every rank, after a device synchronisation)."""
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)
        own = [a for a in self._allocs if a.owner is None]      # torch frees its symmetric-memory buffers itself
        for a in own:
            for r, q in enumerate(a.peers):
                if r != self.rank:
                    self.lib.dfgan_peer_close(C.c_void_p(q))
        self.dist.barrier(group=self.group)
        for a in own:
            self.lib.dfgan_peer_free(C.c_void_p(a.ptr))
        self._allocs = []
