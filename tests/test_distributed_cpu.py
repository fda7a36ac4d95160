"""Host-side sequence-parallel logic on CPU: 2 ranks over gloo (127.0.0.1).

Covers omni_video_b200/distributed/{util,ulysses}.py — the layout algebra of the Ulysses all-to-all
(reference omnivideo/distributed/util.py:95-161, ulysses.py:8-56).  The attention arithmetic itself
is CUDA-only in the product, so here `ulysses.flash_attention` is replaced by the oracle's CPU
attention: what is under test is that scatter-heads/gather-sequence and its inverse reproduce
single-rank attention exactly, and that the gloo route of `all_to_all_single` matches NCCL's
semantics (recv[j] = rank j's send[my_rank]).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dit as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_flash_attention(q, k, v, k_lens=None, window_size=(-1, -1), **kw):
    out = torch.empty_like(q, dtype=torch.float32)
    for b in range(q.shape[0]):
        kl = None if k_lens is None else int(k_lens[b])
        out[b] = O.attention(q[b].float(), k[b].float(), v[b].float(), kl, emulate=False)
    return out.to(q.dtype)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from omni_video_b200.distributed import ulysses, util
        g = util.init_sequence_parallel_group(world)
        assert util.get_sequence_parallel_size() == world
        assert util.get_sequence_parallel_rank() == rank

        # ---- all_to_all_single semantics
        send = torch.arange(world * 3, dtype=torch.float32).view(world, 3) + 100 * rank
        recv = torch.empty_like(send)
        util.all_to_all_single(recv, send, g)
        for j in range(world):
            assert torch.equal(recv[j], torch.arange(3, dtype=torch.float32) + 3 * rank + 100 * j)

        # ---- all_to_all(scatter heads, gather sequence) and its inverse
        gen = torch.Generator().manual_seed(5)
        B, L, H, D = 2, 12, 4, 128
        full = torch.randn(B, L, H, D, generator=gen)
        s = L // world
        mine = full[:, rank * s:(rank + 1) * s].contiguous()
        a = util.all_to_all(mine, scatter_dim=2, gather_dim=1, group=g)
        hg = H // world
        assert torch.equal(a, full[:, :, rank * hg:(rank + 1) * hg])
        back = util.all_to_all(a, scatter_dim=1, gather_dim=2, group=g)
        assert torch.equal(back, mine)

        # ---- gather_forward
        gathered = util.gather_forward(mine, dim=1, group=g)
        assert torch.equal(gathered, full)

        # ---- distributed_attention == single-rank attention (k_lens masks the padded tail)
        ulysses.flash_attention = _oracle_flash_attention
        q = torch.randn(B, L, H, D, generator=gen)
        k = torch.randn(B, L, H, D, generator=gen)
        v = torch.randn(B, L, H, D, generator=gen)
        k_lens = torch.tensor([L, L - 3])
        ref = _oracle_flash_attention(q, k, v, k_lens=k_lens)
        sl = slice(rank * s, (rank + 1) * s)
        out = ulysses.distributed_attention(q[:, sl].contiguous(), k[:, sl].contiguous(),
                                            v[:, sl].contiguous(), k_lens, sp_group=g)
        assert out.shape == (B, s, H, D)
        err = (out - ref[:, sl]).abs().max().item()
        assert err < 1e-5, err

        # ---- gloo has no peer memory: the engine must be told to use the collective transport
        from omni_video_b200.distributed import peer
        assert peer.get_transport(g, torch.device("cpu"), 8, 256) is None

        # ---- fail-fast head divisibility (reference ulysses.py:33-38)
        bad = torch.randn(1, s, 3, D)
        try:
            ulysses.distributed_attention(bad, bad, bad, None, sp_group=g)
            raised = False
        except ValueError:
            raised = True
        assert raised
        results[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        results[rank] = f"{type(ex).__name__}: {ex}"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_ulysses_host_logic_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def _strip_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from omni_video_b200.modules.vae2_1 import Strip, gather_rows, halo_fill
        strip = Strip(dist.new_group(list(range(world))))
        H = 7                                        # uneven strips: 3 + 4 rows
        full = torch.arange(2 * H * 3 * 4, dtype=torch.float32).view(2, H, 3, 4)   # [T, H, W, C]
        h0, h1 = strip.rows(H)
        assert (h0, h1) == ((0, 3) if rank == 0 else (3, 7))
        mine = full[:, h0:h1].contiguous()
        # activations carry one halo slot above and below their rows; halo_fill fills them in place
        ext = torch.full((2, (h1 - h0) + 2, 3, 4), -7.0)
        ext[:, 1:-1] = mine
        halo_fill(ext, strip, True, True)
        want = torch.cat([full[:, h0 - 1:h0] if h0 > 0 else torch.zeros_like(full[:, :1]), mine,
                          full[:, h1:h1 + 1] if h1 < H else torch.zeros_like(full[:, :1])], dim=1)
        assert torch.equal(ext, want)
        only_bottom = torch.full((2, (h1 - h0) + 2, 3, 4), -7.0)
        only_bottom[:, 1:-1] = mine
        halo_fill(only_bottom, strip, False, True)
        assert torch.equal(only_bottom[:, 1:], want[:, 1:]) and bool((only_bottom[:, 0] == -7.0).all())
        sizes = [strip.rows(H, r)[1] - strip.rows(H, r)[0] for r in range(world)]
        assert torch.equal(gather_rows(mine, strip, 1, sizes), full)
        results[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        results[rank] = f"{type(ex).__name__}: {ex}"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_vae_strip_halo_and_gather_two_ranks_gloo():
    """Host logic of the VAE's multi-GPU strips: halo rows come from the right neighbour, image
    borders get zeros, uneven strips gather back in order."""
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_strip_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
