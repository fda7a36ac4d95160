"""Sequence-parallel parity on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sp_check.py

Every rank runs (a) the single-GPU forward of a small DiT and (b) the Ulysses forward over all ranks
(engine path: GEMM epilogues write the all-to-all layouts; and the reference-style per-module hooks
`sp_attn_forward` / `sp_cross_attn_forward` / `sp_dit_forward`), and checks (b) against (a).
"""
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200.distributed.full_sequence_parallel import (sp_attn_forward,  # noqa: E402
                                                                sp_cross_attn_forward, sp_dit_forward)
from omni_video_b200.distributed.util import init_sequence_parallel_group  # noqa: E402
from omni_video_b200.modules.model import WanModel  # noqa: E402


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    heads = 8 if world in (2, 4, 8) else 2 * world
    cfg = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=64, in_dim=16, dim=128 * heads,
               ffn_dim=1024, freq_dim=256, text_dim=256, out_dim=16, num_heads=heads, num_layers=2,
               window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6)
    torch.manual_seed(1818)  # same weights on every rank
    model = WanModel(**cfg).to(dev)
    torch.nn.init.normal_(model.head.head.weight, std=0.02)
    for p in model.parameters():
        if p.dim() == 1:
            p.data.add_(torch.randn_like(p) * 0.02)
    model = model.to(torch.bfloat16).eval().requires_grad_(False)
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = [torch.randn(16, 3, 16, 22, generator=gen, device=dev)]           # 3*8*11 = 264 tokens
    ctx = [torch.randn(40, 256, generator=gen, device=dev).to(torch.bfloat16)]
    t = torch.tensor([500], device=dev)
    L = 264
    seq_len = ((L + world - 1) // world) * world + world                   # forces a padded tail
    ref = model(x, t=t, context=ctx, seq_len=seq_len)[0]

    g = init_sequence_parallel_group(world)
    model.sp_group = g
    # (b1) peer transport: kernels store into the peers' buffers, flag barriers (distributed/peer.py)
    from omni_video_b200.distributed import peer
    os.environ["OVB_SP_TRANSPORT"] = "peer"
    for _ in range(3):                                                     # epochs advance, buffers are reused
        out_engine = model(x, t=t, context=ctx, seq_len=seq_len)[0]
    tr = next((v for v in peer._transports.values() if v), None)
    backend = tr.backend if tr else "unavailable -> NCCL"
    if tr:
        tr.check()
    e1 = rel_l2(out_engine, ref)
    # (b1') the same through a captured CUDA graph (peer barriers keep their epoch on the device)
    model.use_cuda_graph = True
    for _ in range(3):
        out_graph = model(x, t=t, context=ctx, seq_len=seq_len)[0]
    model.use_cuda_graph = False
    graph_same = bool(torch.equal(out_graph, out_engine))
    if tr:
        tr.check()
    # (b2) NCCL all-to-all transport: must give the same bits
    os.environ["OVB_SP_TRANSPORT"] = "nccl"
    out_nccl = model(x, t=t, context=ctx, seq_len=seq_len)[0]
    same = bool(torch.equal(out_nccl, out_engine))
    os.environ["OVB_SP_TRANSPORT"] = "peer"

    for blk in model.blocks:
        blk.self_attn.sp_group = g
        blk.cross_attn.sp_group = g
        blk.self_attn.forward = types.MethodType(sp_attn_forward, blk.self_attn)
        blk.cross_attn.forward = types.MethodType(sp_cross_attn_forward, blk.cross_attn)
    model.forward = types.MethodType(sp_dit_forward, model)
    out_hooks = model(x, t=t, context=ctx, seq_len=seq_len)[0]
    e2 = rel_l2(out_hooks, ref)
    torch.cuda.synchronize()
    print(f"[rank {rank}/{world}] ulysses engine ({backend}) vs single-GPU rel-L2 {e1:.3e}; peer == nccl bits: {same}; graph == eager bits: {graph_same}; "
          f"reference-style hooks {e2:.3e}", flush=True)
    ok = e1 <= 2e-3 and e2 <= 5e-3 and same and graph_same and bool(torch.isfinite(out_engine).all())
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
