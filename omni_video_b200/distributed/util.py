"""Sequence-parallel process-group plumbing — same names and semantics as the reference's
omnivideo/distributed/util.py (:16 init_sequence_parallel_group, :95-121 AllToAll/all_to_all,
:133-161 GatherForward/gather_forward).  torch.distributed (NCCL over NVLink on the B200 box,
gloo in the CPU tests) is plumbing here; the compute either side is libovb200.

Differences that matter on NVSwitch: `all_to_all` moves ONE contiguous buffer with
`all_to_all_single` instead of p chunk tensors + `torch.cat`, and DiTEngine avoids even the
pack/unpack copies by having the GEMM epilogue write the send layout directly (ovb_gemm_bf16
n_split / k_split).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_SEQUENCE_PARALLEL_GROUP = None
_SEQUENCE_PARALLEL_SIZE = 1


def init_distributed_group(backend: str = "nccl"):
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)


def init_sequence_parallel_group(sp_size: int):
    """Consecutive ranks form one SP group (reference util.py:16-40)."""
    global _SEQUENCE_PARALLEL_GROUP, _SEQUENCE_PARALLEL_SIZE
    if not dist.is_initialized():
        return None
    world, rank = dist.get_world_size(), dist.get_rank()
    assert world % sp_size == 0
    mine = None
    for i in range(world // sp_size):
        ranks = list(range(i * sp_size, (i + 1) * sp_size))
        g = dist.new_group(ranks)
        if rank in ranks:
            mine = g
    _SEQUENCE_PARALLEL_GROUP, _SEQUENCE_PARALLEL_SIZE = mine, sp_size
    return mine


def get_sequence_parallel_group():
    return _SEQUENCE_PARALLEL_GROUP


def get_sequence_parallel_size():
    return _SEQUENCE_PARALLEL_SIZE


def get_sequence_parallel_rank():
    return 0 if _SEQUENCE_PARALLEL_GROUP is None else dist.get_rank(_SEQUENCE_PARALLEL_GROUP)


def get_rank(group=None):
    return dist.get_rank(group=_SEQUENCE_PARALLEL_GROUP if group is None else group)


def get_world_size(group=None):
    return dist.get_world_size(group=_SEQUENCE_PARALLEL_GROUP if group is None else group)


def all_to_all_single(recv: torch.Tensor, send: torch.Tensor, group=None):
    """recv[j] <- rank j's send[my_rank]; both [p, ...] contiguous.  NCCL: one grouped send/recv.
    gloo has no all-to-all, so the CPU tests take the all_gather route."""
    p = get_world_size(group)
    assert send.is_contiguous() and recv.is_contiguous() and send.shape[0] == p
    if dist.get_backend(group) == "gloo":
        bucket = [torch.empty_like(send) for _ in range(p)]
        dist.all_gather(bucket, send, group=group)
        r = get_rank(group)
        for j in range(p):
            recv[j].copy_(bucket[j][r])
    else:
        dist.all_to_all_single(recv, send, group=group)
    return recv


def all_to_all(x: torch.Tensor, scatter_dim: int, gather_dim: int, group=None, **kwargs):
    """`scatter` along one dim and `gather` along another (reference util.py:117-121)."""
    p = get_world_size(group)
    if p <= 1:
        return x
    send = torch.stack([u.contiguous() for u in x.chunk(p, dim=scatter_dim)])
    recv = torch.empty_like(send)
    all_to_all_single(recv, send, group)
    return torch.cat(list(recv.unbind(0)), dim=gather_dim).contiguous()


def all_gather(tensor: torch.Tensor, group=None):
    p = get_world_size(group)
    if p == 1:
        return [tensor]
    out = [torch.empty_like(tensor) for _ in range(p)]
    dist.all_gather(out, tensor.contiguous(), group=group)
    return out


def gather_forward(input: torch.Tensor, dim: int, group=None):
    """all-gather along `dim` (reference util.py:157-161; inference only, no autograd)."""
    if get_world_size(group) <= 1:
        return input
    return torch.cat(all_gather(input, group), dim=dim).contiguous()
