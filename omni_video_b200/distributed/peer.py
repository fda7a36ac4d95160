"""Peer-memory transport for Ulysses sequence parallelism on one NVSwitch node.

The reference's `distributed_attention` (omnivideo/distributed/ulysses.py:41-55) is
all-to-all -> attention -> all-to-all.  Here both exchanges disappear into the kernels either side:

    q, k : ovb_rmsnorm_rope (scatter form)   normalises/rotates a row and stores each rank's head
    v    : ovb_rmsnorm_rope (copy_only)      group straight into THAT rank's receive buffer (NVLink
                                             stores from the kernel; no send buffer round trip)
    out  : ovb_attn_fwd (scatter form)       the attention epilogue stores every query row into the
                                             receive buffer of the rank that owns the row
    sync : ovb_peer_signal_wait              one 32-thread launch per exchange: system fence, flag to
                                             every peer, wait for every peer's flag

so data moves tile by tile while the producing kernel is still computing, and NCCL is not on the
block's critical path at all.  The receive buffers and the flag pads are symmetric allocations that
every rank of the group has mapped: `torch.distributed._symmetric_memory` (CUDA VMM handles) when it
works, otherwise CUDA IPC handles of ordinary allocations exchanged through the process group.

Write-after-read safety with single buffers: a rank writes block i+1's q/k/v into a peer only after
passing the "out" barrier of block i, which that peer signals after its attention (the reader of
q/k/v) of block i; and it writes block i+1's attention output only after the "qkv" barrier of block
i+1, which the peer signals after its O-projection (the reader of the output buffer) of block i.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import _lib, ops

bf16 = torch.bfloat16


class PeerUnavailable(RuntimeError):
    pass


def _enable_peer_access(devices):
    lib = _lib.load()
    for d in devices:
        _lib.check(lib.ovb_enable_peer_access(int(d)), "ovb_enable_peer_access")


class _Symmetric:
    """`nbytes` of device memory on every rank of `group`, mapped into every rank.

    Construction is split in two so that a failure on ONE rank can never leave the others inside a
    collective: `allocate()` is purely local (allocate + export), `exchange()` holds the collectives
    (rendezvous / all_gather_object) and is entered only after every rank has reported a successful
    `allocate()` (see `_make_symmetric`)."""

    def __init__(self, nbytes: int, device: torch.device, group, how: str):
        if how not in ("symm_mem", "ipc"):
            raise ValueError(how)
        self.nbytes, self.device, self.group, self.how = nbytes, device, group, how
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def allocate(self):
        if self.how == "symm_mem":
            import torch.distributed._symmetric_memory as symm_mem
            self._symm_mem = symm_mem
            self.local = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        else:
            # an ordinary cudaMalloc block of its own (not a slice of a cached segment shared with
            # other tensors), exported as a CUDA IPC handle
            self.local = torch.empty(max(self.nbytes, 2 << 20), dtype=torch.uint8, device=self.device)
            self._meta = self.local.untyped_storage()._share_cuda_()

    def exchange(self):
        """Collective part.  Returns None or the exception of the local, post-collective step (opening
        the peers' handles), which the caller again agrees on across ranks."""
        if self.how == "symm_mem":
            self._hdl = self._symm_mem.rendezvous(self.local, group=self.group)
            self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        else:
            metas = [None] * self.world
            dist.all_gather_object(metas, self._meta, group=self.group)
            try:
                self._keep, self.ptrs = [], []
                for r, m in enumerate(metas):
                    if r == self.rank:
                        self.ptrs.append(self.local.data_ptr())
                        continue
                    st = torch.UntypedStorage._new_shared_cuda(*m)
                    self._keep.append(st)
                    self.ptrs.append(st.data_ptr())
                _enable_peer_access({m[0] for r, m in enumerate(metas) if r != self.rank})
            except Exception as ex:  # noqa: BLE001
                return ex
        if self.ptrs[self.rank] != self.local.data_ptr():
            return RuntimeError("own mapping differs from the local pointer")
        return None


def _all_ok(ok: bool, device, group) -> bool:
    flag = torch.full((1,), 1 if ok else 0, device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return int(flag.item()) == 1


def _make_symmetric(nbytes, device, group) -> _Symmetric:
    order = os.environ.get("OVB_PEER_BACKEND", "symm_mem,ipc").split(",")
    errs = []
    for how in order:
        how = how.strip()
        s, err = None, None
        try:                                    # local phase: no collective inside
            s = _Symmetric(nbytes, device, group, how)
            s.allocate()
        except Exception as ex:  # noqa: BLE001 - every rank must agree on the outcome below
            err = ex
        if not _all_ok(err is None, device, group):
            errs.append(f"{how} (allocate): {type(err).__name__ if err else 'peer rank'}: {err}")
            continue
        # exchange phase: every rank enters the same collectives; a backend whose collective itself
        # raises on one rank only cannot be recovered from and propagates
        err = s.exchange()
        if _all_ok(err is None, device, group):
            return s
        errs.append(f"{how} (exchange): {type(err).__name__ if err else 'peer rank'}: {err}")
    raise PeerUnavailable("no peer-memory backend worked on every rank: " + " | ".join(errs))


class UlyssesPeerTransport:
    """Receive buffers + flag pad of one sequence-parallel group for a given (rows, width) shape."""

    def __init__(self, group, device: torch.device, rows: int, d: int):
        self.group = group
        self.sp, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.sp <= 8 and d % self.sp == 0
        self.rows, self.d, self.dg = rows, d, d // self.sp
        sp, dg = self.sp, self.dg
        qkv_bytes = sp * rows * 3 * dg * 2
        att_bytes = sp * rows * dg * 2
        pad_off = (qkv_bytes + att_bytes + 255) // 256 * 256
        self._sym = _make_symmetric(pad_off + 256, device, group)
        self.backend = self._sym.how
        base = self._sym.local
        base[pad_off:].zero_()
        self.qkv_recv = base[:qkv_bytes].view(bf16).view(sp, rows, 3 * dg)
        self.att_recv = base[qkv_bytes:qkv_bytes + att_bytes].view(bf16).view(sp, rows, dg)
        me = self.rank
        # where THIS rank's slice lands inside every peer's receive buffers (src-rank-major layout)
        self.qkv_dst = [p + me * rows * 3 * dg * 2 for p in self._sym.ptrs]
        self.att_dst = [p + qkv_bytes + me * rows * dg * 2 for p in self._sym.ptrs]
        self.pad_ptrs = [p + pad_off for p in self._sym.ptrs]
        self.state = torch.zeros(2, device=device, dtype=torch.int32)   # (epoch counter, time-out status)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)            # pads are zero everywhere before the first signal

    def q_k_v_dst(self, which: int) -> List[int]:
        """Destination pointers (one per rank) of the q (0), k (1) or v (2) column block."""
        return [p + which * self.dg * 2 for p in self.qkv_dst]

    def barrier(self):
        ops.peer_signal_wait(self.pad_ptrs, self.rank, self.state)

    def check(self):
        """Raise if any barrier timed out (call outside the timed region: it synchronises).
        A time-out (a rank more than ~30 s behind, or dead) is fatal on the device as well: the barrier
        kernel traps after recording it here, so ranks of a group must stay in lock-step."""
        try:
            bad = int(self.state[1].item()) != 0
        except RuntimeError as ex:              # the trap surfaces as a launch failure at this sync
            raise RuntimeError("peer barrier timed out (device trap): a rank of the sequence-parallel "
                               "group fell behind by more than ~30 s or died") from ex
        if bad:
            raise RuntimeError("peer barrier timed out: a rank of the sequence-parallel group fell behind or died")


_transports = {}


def get_transport(group, device, rows, d) -> Optional[UlyssesPeerTransport]:
    """Cached per (group, shape).  Returns None when OVB_SP_TRANSPORT=nccl or the group is not NCCL."""
    if os.environ.get("OVB_SP_TRANSPORT", "peer") != "peer" or dist.get_backend(group) != "nccl":
        return None
    key = (id(group), device.index, rows, d)
    t = _transports.get(key)
    if t is None:
        try:
            t = UlyssesPeerTransport(group, device, rows, d)
        except PeerUnavailable as ex:
            import warnings
            warnings.warn(f"peer transport unavailable, using NCCL all-to-all: {ex}")
            t = False
        _transports[key] = t
    return t or None
