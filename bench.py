"""bench.py — DiT denoise steps/s on synthetic latents (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model 1.3B|A14B] [--video 480p|720p] [--text-len 512]

One *step* is what the reference's sampling loop does per timestep (x2x_gen_unified.py:474-492):
two `WanModel.forward` calls (conditional + unconditional, B=1 each), the classifier-free-guidance
combine and the flow-matching UniPC update of the latent (utils/fm_solvers_unipc.py).  At N=1 the workload is BASELINE.json `configs[1]`:
OmniVideo2-1.3B, 81x480x832 video = latent [16,21,60,104] = 32760 tokens, bf16, random-init weights
of that architecture, synthetic N(0,1) latents / text context (no checkpoints or datasets offline).

What is printed (ONE JSON line, rank 0):
  value        steps/s with the latents and contexts already resident in HBM (device-timed, CUDA
               events, max over ranks);
  e2e          the same step through the public drop-in API (`omni_video_b200...WanModel.forward`)
               with HOST buffers: every step copies latents + both contexts from pinned host memory
               and reads the updated latent back, all inside the timed region;
  roofline     the dominant kernel (attn_fwd_kernel, tcgen05 flash attention): algorithmic FLOPs of
               all its launches in the timed region / their CUDA-event durations, against the
               measured sustained bf16 peak of MEASURED_PEAKS.json;
  cpu_baseline the oracle port of the reference path (oracle/dit.py) timed on this box's host
               cores on a bounded sample of the same workload.

N>1 (torchrun), total work fixed => "scaling": "strong".  The step has two independent forwards
(cond / uncond), so an even N first splits them over two groups of N/2 ranks that exchange the
8.4 MB noise prediction once per step (no collective inside a forward); inside a
group the token dimension is sharded with Ulysses all-to-all (omni_video_b200/distributed).
N=2 -> cfg2, N=4 -> cfg2 x ulysses2, N=8 -> cfg2 x ulysses4 (12 heads % 8 != 0 rules out
ulysses8 for the 1.3B model anyway).  `--no-cfg-split` gives the reference's layout (one Ulysses
group of N ranks running both forwards).

`--impl reference`: the reference's own implementation of this path is Python/PyTorch and ships
no kernels; its CPU path (restated in oracle/dit.py and pinned to the reference's outputs by
tests/test_oracle_golden.py) is timed on the host cores with all threads, each step a bounded
sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = {
    # reference omnivideo/configs/wan_t2v_1_3B.py:19-40, wan_t2v_A14B.py:19-42
    "1.3B": dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30),
    "A14B": dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40),
}
VIDEOS = {"480p": (21, 60, 104), "720p": (21, 90, 160), "256": (5, 32, 32)}
GUIDE_SCALE = 3.0  # tools/inference/inference_omni_t2v.sh:43-53


def model_cfg(name: str, text_len: int) -> dict:
    cfg = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=text_len, in_dim=16, freq_dim=256,
               text_dim=4096, out_dim=16, window_size=(-1, -1), qk_norm=True, cross_attn_norm=True,
               eps=1e-6)
    cfg.update(MODELS[name])
    return cfg


def forward_flops(cfg: dict, L: int, Lc: int) -> dict:
    """ALGORITHMIC FLOPs of one forward."""
    d, f, N = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    attn = N * (4.0 * L * L * d + 4.0 * L * Lc * d)
    gemm = N * (8.0 * L * d * d + 4.0 * L * d * d + 4.0 * Lc * d * d + 4.0 * L * d * f)
    other = 2.0 * 64 * L * d + 2.0 * Lc * 4096 * d + 2.0 * Lc * d * d + (2 * 256 * d + 14 * d * d) + 2.0 * L * d * 64
    return dict(total=attn + gemm + other, attn=attn, gemm=gemm)


def load_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            j = json.load(fh)
        return dict(tflops=float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0))),
                    tflops_burst=float(j.get("bf16_tflops", 1590.0)), hbm=float(j.get("hbm_gbs", 6650.0)),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc, self.path = None, None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        # samples under load = those drawing well above idle power
        hot = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        hot.sort()
        return dict(sm_mhz=hot[len(hot) // 2], sm_max_mhz=max(mx), power_w_max=max(pw),
                    samples=len(sm), reasons=sorted(reasons))


# --------------------------------------------------------------------------------- our arm
def init_random_(model, seed: int):
    """Random-init weights of the architecture, generated on the device (no checkpoints offline).
    Same scheme as the reference's init_weights (model.py:551-573) except that head.head.weight is
    not zero and biases/modulation get small non-zero values."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and "modulation" not in name:
                fan_out, fan_in = p.shape[0], p[0].numel()
                a = math.sqrt(6.0 / (fan_in + fan_out))
                if name.startswith(("text_embedding", "time_embedding")):
                    p.copy_(torch.randn(p.shape, generator=g, device="cuda") * 0.02)
                else:
                    p.copy_((torch.rand(p.shape, generator=g, device="cuda") * 2 - 1) * a)
            elif "modulation" in name:
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") / math.sqrt(p.shape[-1]))
            elif name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * 0.02)
            else:  # norm weights
                p.copy_(1.0 + torch.randn(p.shape, generator=g, device="cuda") * 0.05)


class Stepper:
    """The reference's per-timestep work (x2x_gen_unified.py:474-492) on our WanModel."""

    def __init__(self, model, seq_len, world, rank, cfg_groups, batch_cfg=False):
        import torch
        self.batch_cfg = batch_cfg
        self.torch, self.model, self.seq_len = torch, model, seq_len
        self.world, self.rank = world, rank
        self.cfg_groups = cfg_groups  # None, or (my_branch, pair_group) when cond/uncond are split
        # the reference's sampler: flow-matching UniPC, 50 steps, shift 5 (x2x_gen_unified.py:409-416,
        # tools/inference/inference_omni_t2v.sh:43-53); ours runs each tensor expression as one launch
        from omni_video_b200.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
        self.sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        self.sched_pos = 50

    def step(self, latent, ctx_c, ctx_u, i):
        torch = self.torch
        if self.sched_pos >= 50:  # a fresh 50-step schedule (the bench runs fewer steps than one sample needs)
            self.sched.set_timesteps(50, device=latent.device, shift=5.0)
            self.sched_pos = 0
        t = self.sched.timesteps[self.sched_pos:self.sched_pos + 1]
        self.sched_pos += 1
        if self.cfg_groups is None and self.batch_cfg:
            # cond + uncond as ONE forward with B = 2 (WanModel.forward takes lists): same
            # arithmetic per sample, half the launches, and 2x the CTAs per attention launch
            # (12 heads x 128 query blocks = 10.4 waves of 148 SMs -> 20.8: less tail waste)
            cond, uncond = self.model([latent, latent], t=torch.cat([t, t]), context=[ctx_c, ctx_u],
                                      seq_len=self.seq_len)
        elif self.cfg_groups is None:
            cond = self.model([latent], t=t, context=[ctx_c], seq_len=self.seq_len)[0]
            uncond = self.model([latent], t=t, context=[ctx_u], seq_len=self.seq_len)[0]
        else:
            import torch.distributed as dist
            branch, pair = self.cfg_groups
            mine = self.model([latent], t=t, context=[ctx_c if branch == 0 else ctx_u],
                              seq_len=self.seq_len)[0]
            both = torch.empty((2,) + tuple(mine.shape), device=mine.device, dtype=mine.dtype)
            dist.all_gather_into_tensor(both, mine.contiguous(), group=pair)
            cond, uncond = both[0], both[1]
        # noise_pred = uncond + g (cond - uncond) (:484) folded into the scheduler's x0 prediction;
        # UniC corrector + UniP predictor (:486-491)
        return self.sched.step_cfg(cond, uncond, GUIDE_SCALE, t[0], latent.unsqueeze(0))[0].squeeze(0)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from omni_video_b200 import _lib, ops
    from omni_video_b200.modules.model import WanModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    _lib.check(lib.ovb_check_device(), "ovb_check_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = model_cfg(args.model, args.text_len)
    Fr, H, W = VIDEOS[args.video]
    L = Fr * (H // 2) * (W // 2)
    heads = cfg["num_heads"]

    # ---- parallel layout
    sp, cfg_par = world, 1
    if world > 1 and (heads % world != 0 or args.cfg_split or (world % 2 == 0 and not args.no_cfg_split)):
        assert world % 2 == 0 and heads % (world // 2) == 0, f"{heads} heads cannot be split over {world} ranks"
        sp, cfg_par = world // 2, 2
    sp_group, cfg_groups = None, None
    if world > 1:
        for g0 in range(0, world, sp):
            ranks = list(range(g0, g0 + sp))
            g = dist.new_group(ranks)
            if rank in ranks:
                sp_group = g
        if cfg_par == 2:
            for r0 in range(sp):
                ranks = [r0, r0 + sp]
                g = dist.new_group(ranks)
                if rank in ranks:
                    cfg_groups = (rank // sp, g)
    seq_len = int(math.ceil(L / sp)) * sp  # x2x_gen_unified.py:332-334

    torch.manual_seed(1818)
    with torch.device("cuda"):
        model = WanModel(**cfg, skip_init=True)
    model = model.to(torch.bfloat16).eval().requires_grad_(False)
    init_random_(model, 1818)
    if sp_group is not None and sp > 1:
        model.sp_group = sp_group
    # CUDA-graph replay of the forward: default whenever sequence parallelism cuts the kernels short
    # (N >= 4 here); the captured graph holds this library's kernels only (no NCCL op: the final
    # gather runs eagerly after every replay), see --graph / --no-graph
    use_graph = bool(args.graph) if args.graph is not None else (sp > 1)
    model.use_cuda_graph = False

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1818)
    latent0 = torch.randn((16, Fr, H, W), generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)  # bf16 latents as x2x_gen_unified.py:382-391
    ctx_c = torch.randn((min(512, args.text_len), 4096), generator=gen, device=dev).to(torch.bfloat16)
    ctx_u = torch.randn((min(512, args.text_len), 4096), generator=gen, device=dev).to(torch.bfloat16)
    stepper = Stepper(model, seq_len, world, rank, cfg_groups,
                      batch_cfg=(sp == 1 and cfg_par == 1 and not args.no_cfg_batch))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- parity of the parallel layout, before anything is timed ----------------
    # The same forward (conditional branch, step-0 inputs) on the sequence-parallel layout, with the
    # peer-memory transport and with the NCCL all-to-all transport, against the SINGLE-GPU forward of
    # the same inputs (every rank holds the full weights; reference semantics:
    # distributed/full_sequence_parallel.py:123-126,163 and ulysses.py:41-55 are re-orderings of the
    # single-GPU computation).
    parity = None
    if world > 1:
        t_par = torch.full((1,), 999, device=dev, dtype=torch.int64)
        if sp > 1:
            from omni_video_b200.distributed import peer as peer_mod
            o_sp = model([latent0], t=t_par, context=[ctx_c], seq_len=seq_len)[0].clone()
            tr = next((v for v in peer_mod._transports.values() if v), None)
            if tr is not None:
                tr.check()
            prev = os.environ.get("OVB_SP_TRANSPORT")
            os.environ["OVB_SP_TRANSPORT"] = "nccl"
            o_nccl = model([latent0], t=t_par, context=[ctx_c], seq_len=seq_len)[0].clone()
            if prev is None:
                os.environ.pop("OVB_SP_TRANSPORT")
            else:
                os.environ["OVB_SP_TRANSPORT"] = prev
            model.sp_group = None
            o_one = model([latent0], t=t_par, context=[ctx_c], seq_len=seq_len)[0]
            model.sp_group = sp_group
            err = float((o_sp.double() - o_one.double()).norm() / o_one.double().norm())
            stats = torch.tensor([err, 0.0 if torch.equal(o_sp, o_nccl) else 1.0,
                                  0.0 if bool(torch.isfinite(o_sp).all()) else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
            parity = dict(sp_vs_single_rel_l2=float(f"{stats[0].item():.3e}"),
                          peer_vs_nccl_equal=bool(stats[1].item() == 0.0), finite=bool(stats[2].item() == 0.0),
                          transport=(tr.backend if tr is not None else "nccl"),
                          what=f"one forward, {L} tokens, ulysses{sp} vs the same forward on one GPU, max over ranks")
            del o_sp, o_nccl, o_one
        else:
            parity = dict(sp_vs_single_rel_l2=0.0, peer_vs_nccl_equal=None, finite=True, transport=None,
                          what="cfg2 only: every forward runs whole on one GPU (nothing is sharded)")
        torch.cuda.empty_cache()
    model.use_cuda_graph = use_graph

    # ---------------- device-resident run (value) ----------------
    latent = latent0.clone()
    # the sampler starts before the warm-up steps (nvidia-smi needs ~0.3 s to produce its first
    # line, longer than a multi-GPU timed region); only samples drawing >= half of the peak power
    # seen, i.e. taken under load, enter the median
    clocks = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        latent = stepper.step(latent, ctx_c, ctx_u, i)
    barrier()
    # per-launch CUDA events (roofline leg) need individual launches: in graph mode they are taken
    # on one extra eager step right after the timed region instead of inside it
    ops.timeline = None if use_graph else []
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.cudart().cudaProfilerStart()  # `ncu --profile-from-start off` sees only the timed steps
    e0.record()
    for i in range(args.steps):
        latent = stepper.step(latent, ctx_c, ctx_u, args.warmup + i)
    e1.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - n0
    timeline_steps = args.steps
    if use_graph:
        model.use_cuda_graph = False
        ops.timeline = []
        latent = stepper.step(latent, ctx_c, ctx_u, args.warmup + args.steps)
        barrier()
        model.use_cuda_graph = True
        timeline_steps = 1
    timeline, ops.timeline = ops.timeline, None
    clk = clocks.stop() if clocks is not None else None
    finite = bool(torch.isfinite(latent).all().item())

    # ---------------- end-to-end run through the public API with host buffers (e2e) ------------
    h_lat = latent0.cpu().pin_memory()
    h_c, h_u = ctx_c.cpu().pin_memory(), ctx_u.cpu().pin_memory()
    h_out = torch.empty_like(h_lat).pin_memory()
    h2d = h_lat.numel() * h_lat.element_size() + (h_c.numel() + h_u.numel()) * h_c.element_size()
    d2h = h_out.numel() * h_out.element_size()

    def e2e_step(i):
        lat = h_lat.to(dev, non_blocking=True)
        c, u = h_c.to(dev, non_blocking=True), h_u.to(dev, non_blocking=True)
        new = stepper.step(lat, c, u, i)
        h_out.copy_(new, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller owns the result only after this
        h_lat.copy_(h_out)

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        e2e_step(i)
    g1.record()
    barrier()
    e2e_ms = max_over_ranks(max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3))

    # ---------------- VAE decode of the sampled latent (all ranks: H-strips with halo exchange) ---
    peaks = load_peaks()
    vae = None
    if not args.no_vae:
        vae = vae_decode_leg((Fr, H, W), dev, peaks, dist.group.WORLD if world > 1 else None, world)

    def shutdown():
        """Leave the job: drop the captured graphs, then tear NCCL down.  (Round 1 captured the final
        all-gather inside the graph and destroy_process_group() then waited for ever; the capture now
        ends before the gather.  A watchdog still guarantees that a teardown problem cannot hang the
        job after the result line has been printed.)"""
        if world == 1:
            return
        torch.cuda.synchronize()
        dist.barrier()                          # leave together: rank 0 prints its line before anyone goes
        torch.cuda.synchronize()
        if use_graph and getattr(model, "_engine", None) is not None:
            model._engine._graphs.clear()
        sys.stdout.flush()
        sys.stderr.flush()
        import threading

        def _bail():
            sys.stderr.write(f"[rank {rank}] destroy_process_group did not return within 30 s; leaving\n")
            sys.stderr.flush()
            os._exit(0)
        timer = threading.Timer(30.0, _bail)
        timer.daemon = True
        timer.start()
        dist.destroy_process_group()
        timer.cancel()

    if rank != 0:
        shutdown()
        return

    # ---------------- roofline of the dominant kernel ----------------
    torch.cuda.synchronize()
    agg = {}
    for kind, flops, shape, s, e in timeline:
        a = agg.setdefault(kind, [0.0, 0.0, 0])
        a[0] += flops; a[1] += s.elapsed_time(e); a[2] += 1
    attn_tf = agg["attn"][0] / agg["attn"][1] / 1e9 if "attn" in agg else 0.0
    gemm_tf = agg["gemm"][0] / agg["gemm"][1] / 1e9 if "gemm" in agg else 0.0
    fl = forward_flops(cfg, L, args.text_len)
    steps_per_s = args.steps / (dev_ms / 1e3)
    whole_tf = 2.0 * fl["total"] * steps_per_s / 1e12
    traffic = None
    try:  # per-launch DRAM bytes of the self-attention launch from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)["attn_fwd_kernel"]
        b_timed = 2 if stepper.batch_cfg else 1
        key = f"B{b_timed}_H{heads // sp}_Lq{seq_len}_Lk{seq_len}"
        traffic = tj.get(key)
    except Exception:
        traffic = None
    roofline = dict(bound="tensor", kernel="attn_fwd_kernel (self+cross flash attention, tcgen05)",
                    achieved=round(attn_tf, 1), peak=peaks["tflops"], unit="TFLOP/s",
                    frac=round(attn_tf / peaks["tflops"], 4), traffic=traffic,
                    traffic_note=(f"DRAM bytes of one self-attention launch of the timed shape "
                                  f"(B={2 if stepper.batch_cfg else 1}, {heads // sp} heads, {seq_len} tokens; "
                                  "ncu --set full, profiles/ncu_traffic.json); algorithmic q+k+v+o bytes = "
                                  f"{(2 if stepper.batch_cfg else 1) * 4 * seq_len * (heads // sp) * 128 * 2 / 1e6:.1f} MB"),
                    peak_source=peaks["source"],
                    launches=agg.get("attn", [0, 0, 0])[2],
                    ms_per_launch=round(agg["attn"][1] / max(1, agg["attn"][2]), 4) if "attn" in agg else None,
                    share_of_step=round(agg["attn"][1] / (dev_ms * timeline_steps / args.steps), 4) if "attn" in agg else None,
                    timed_from=("one extra eager step after the timed region (the timed steps replay a CUDA graph)"
                                if use_graph else "CUDA events around every launch inside the timed region"),
                    gemm_kernel=dict(achieved=round(gemm_tf, 1), frac=round(gemm_tf / peaks["tflops"], 4),
                                     launches=agg.get("gemm", [0, 0, 0])[2],
                                     share_of_step=round(agg["gemm"][1] / (dev_ms * timeline_steps / args.steps), 4) if "gemm" in agg else None),
                    whole_step=dict(achieved=round(whole_tf / world, 1),
                                    frac=round(whole_tf / world / peaks["tflops"], 4),
                                    tflop_per_step=round(2.0 * fl["total"] / 1e12, 1)))

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_leg(cfg, (Fr, H, W), args.text_len, budget_s=args.cpu_budget)

    line = dict(metric="DiT denoise steps/s", value=round(steps_per_s, 5), unit="steps/s",
                n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=round(dev_ms / args.steps, 3), higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=f"OmniVideo2-{args.model} DiT, {args.video} "
                                     f"(latent 16x{Fr}x{H}x{W}, {L} tokens), text_len {args.text_len}, "
                                     "step = cond+uncond forward + CFG + UniPC update, random-init weights",
                            cfg_batched=bool(stepper.batch_cfg), cuda_graph=bool(use_graph),
                            sp_transport=(os.environ.get("OVB_SP_TRANSPORT", "peer") if sp > 1 else None),
                            parallelism=(f"ulysses{sp}" + (f"xcfg{cfg_par}" if cfg_par > 1 else "")) if world > 1 else "single",
                            l2="inputs+weights per step (>2.8 GB) exceed the 126 MB L2; no flush needed",
                            forwards_per_s=round(2 * steps_per_s, 5), finite=finite),
                e2e=dict(value=round(args.steps / (e2e_ms / 1e3), 5), unit="steps/s",
                         h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                gpu_launches=int(launches), clocks=clk, roofline=roofline)
    if parity is not None:
        line["parity"] = parity
    if vae is not None:
        line["vae_decode"] = vae
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    shutdown()


def vae_decode_leg(grid_lat, dev, peaks, group=None, world=1):
    """BASELINE.json configs[1] ends with one Wan2.1 VAE decode of the sampled latent (the reference
    decodes on rank 0 only, x2x_gen_unified.py:512-524; with N > 1 we split the clip into H-strips
    over all ranks, halo rows exchanged before every spatial conv).  Reported beside the step rate,
    not folded into it: one decode per 50 steps.  Algorithmic FLOPs: SURVEY.md Appendix A cost model (measured on the reference with
    FlopCounterMode): 0.688 GFLOP per latent pixel for frame 0 + 2.162 per later frame, plus the
    per-frame single-head attention."""
    import torch
    from omni_video_b200 import ops
    from omni_video_b200.modules.vae2_1 import Wan2_1_VAE
    torch.cuda.empty_cache()
    T, h, w = grid_lat
    vae = Wan2_1_VAE(vae_pth=None, device=dev)
    g = torch.Generator(device="cuda").manual_seed(1818)
    for name, p in vae.model.named_parameters():
        if p.dim() > 1 and "gamma" not in name:
            p.data.copy_(torch.randn(p.shape, generator=g, device="cuda") / p[0].numel() ** 0.5)
    z = torch.randn((16, T, h, w), generator=g, device=dev)
    if group is not None:
        vae.set_strip_group(group)
    flops = (0.688e9 + 2.162e9 * (T - 1)) * h * w + 4.0 * (h * w) ** 2 * 384 * T
    times = []
    for it in range(3):
        ops.timeline = [] if it == 2 else None
        if group is not None:
            import torch.distributed as dist
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = vae.decode([z])[0]
        e1.record()
        torch.cuda.synchronize()
        ms_it = e0.elapsed_time(e1)
        if group is not None:
            tt = torch.tensor([ms_it], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_it = float(tt.item())
        times.append(ms_it)
    tl, ops.timeline = ops.timeline, None
    strips_vs_single = None
    if group is not None:
        # parity of the strip layout: the same decode whole on this GPU (bf16 kernels in a different
        # tiling order: equal up to accumulation-order rounding)
        vae.set_strip_group(None)
        whole = vae.decode([z])[0]
        d = (out - whole).abs().max()
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        mse = float(((out.double() - whole.double()) ** 2).mean().item())
        strips_vs_single = dict(max_abs=float(f"{d.item():.3e}"),
                                psnr_db=round(10.0 * math.log10(4.0 / max(mse, 1e-30)), 1))
        del whole
    conv_fl = sum(t[1] for t in tl if t[0] == "conv")
    conv_ms = sum(t[3].elapsed_time(t[4]) for t in tl if t[0] == "conv")
    norm_b = sum(t[1] for t in tl if t[0] == "norm")
    norm_ms = sum(t[3].elapsed_time(t[4]) for t in tl if t[0] == "norm")
    ms = min(times[1:])
    res = dict(ms=round(ms, 1), out_shape=list(out.shape), n_gpus=world,
               layout="single GPU" if world == 1 else f"{world} H-strips, halo exchange per spatial conv",
               tflops=round(flops / ms / 1e9, 1),
               frac_of_bf16_peak=round(flops / ms / 1e9 / peaks["tflops"] / world, 4),
               conv_kernel=dict(tflops=round(conv_fl / max(conv_ms, 1e-9) / 1e9, 1), share=round(conv_ms / ms, 3)),
               norm_kernel=dict(gbs=round(norm_b / max(norm_ms, 1e-9) / 1e6, 1), share=round(norm_ms / ms, 3),
                                frac_of_hbm_peak=round(norm_b / max(norm_ms, 1e-9) / 1e6 / peaks["hbm"], 4)),
               finite=bool(torch.isfinite(out).all().item()))
    if strips_vs_single is not None:
        res["parity_strips_vs_single_gpu"] = strips_vs_single
    del out, vae
    torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------- CPU legs
def _cpu_block_seconds(cfg, grid_lat, Lc, threads):
    """One DiT block (of num_layers) of the oracle port on the host, fp32, at the given latent grid."""
    import torch
    from oracle import dit as O
    from oracle import synth
    torch.set_num_threads(threads)
    Fr, H, W = grid_lat
    grid = (Fr, H // 2, W // 2)
    L = grid[0] * grid[1] * grid[2]
    d = cfg["dim"]
    sd = synth.dit_state_dict(cfg, seed=1818, num_layers=1)
    rng = torch.Generator().manual_seed(7)
    x = torch.randn((L, d), generator=rng)
    e0 = torch.randn((6, d), generator=rng) * 0.1
    ctx = torch.randn((Lc, d), generator=rng)
    ang = O.rope_angles(1024, d // cfg["num_heads"])
    t0 = time.perf_counter()
    with torch.no_grad():
        y = O.block_forward(sd, "blocks.0.", x, e0, L, grid, ang, ctx, cfg["num_heads"], cfg["eps"],
                            False, False)
    dt = time.perf_counter() - t0
    assert bool(torch.isfinite(y).all())
    return dt, L


def cpu_reference_leg(cfg, grid_lat, Lc, budget_s=30.0):
    """Bounded sample of the workload on the host cores: ONE of the num_layers DiT blocks at the
    full token count when that fits the time budget, otherwise at fewer latent frames and scaled
    by the algorithmic-FLOP ratio.  Extrapolated to steps/s = 1 / (2 forwards x N blocks x t)."""
    threads = os.cpu_count() or 1
    Fr, H, W = grid_lat
    N = cfg["num_layers"]

    def blk_flops(L):
        d, f = cfg["dim"], cfg["ffn_dim"]
        return 4.0 * L * L * d + 4.0 * L * Lc * d + 12.0 * L * d * d + 4.0 * Lc * d * d + 4.0 * L * d * f

    cal_t, cal_L = _cpu_block_seconds(cfg, (1, H, W), Lc, threads)           # 1 latent frame
    rate = blk_flops(cal_L) / cal_t
    full_L = Fr * (H // 2) * (W // 2)
    frames = Fr
    while frames > 1 and 2.0 * blk_flops(frames * (H // 2) * (W // 2)) / rate > budget_s:
        frames -= 1
    t, Ls = _cpu_block_seconds(cfg, (frames, H, W), Lc, threads)
    t_full_block = t * blk_flops(full_L) / blk_flops(Ls)
    steps_per_s = 1.0 / (2.0 * N * t_full_block)
    sample = (f"1 of {N} DiT blocks, fp32 oracle port (oracle/dit.py), {frames}/{Fr} latent frames "
              f"({Ls} of {full_L} tokens) in {t:.1f} s" +
              ("" if frames == Fr else ", scaled to the full token count by algorithmic FLOPs") +
              f"; extrapolated x{N} blocks x2 forwards")
    return dict(value=float(f"{steps_per_s:.3e}"), unit="steps/s", cores=threads, kind="port", sample=sample,
                cpu_tflops=round(blk_flops(Ls) / t / 1e12, 3))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_cfg(args.model, args.text_len)
    Fr, H, W = VIDEOS[args.video]
    L = Fr * (H // 2) * (W // 2)
    budget = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_leg(cfg, (Fr, H, W), args.text_len, budget_s=budget)
    vals = []
    last = None
    for _ in range(args.steps):
        last = cpu_reference_leg(cfg, (Fr, H, W), args.text_len, budget_s=budget)
        vals.append(last["value"])
    v = len(vals) / sum(1.0 / x for x in vals)  # steps / total extrapolated seconds
    last["value"] = float(f"{v:.3e}")
    line = dict(impl="reference", metric="DiT denoise steps/s", value=last["value"], unit="steps/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=round(1e3 / v, 1),
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"OmniVideo2-{args.model} DiT, {args.video} (latent 16x{Fr}x{H}x{W}, "
                                     f"{L} tokens), text_len {args.text_len}, reference CPU path "
                                     "(PyTorch fp32 restatement, host cores), EXTRAPOLATED from a bounded "
                                     "sample: see cpu_baseline.sample", extrapolated=True),
                e2e=dict(value=last["value"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, cpu_baseline=last)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="1.3B", choices=sorted(MODELS))
    ap.add_argument("--video", default="480p", choices=sorted(VIDEOS))
    ap.add_argument("--text-len", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true", help="skip the VAE decode leg")
    ap.add_argument("--no-cfg-batch", action="store_true",
                    help="N=1: run cond and uncond as two B=1 forwards (the reference's loop) instead of one B=2 forward")
    ap.add_argument("--no-cfg-split", action="store_true",
                    help="N>1: one Ulysses group of N ranks running cond then uncond (the reference's layout) "
                         "instead of the default cond/uncond split over two groups of N/2")
    ap.add_argument("--cfg-split", action="store_true",
                    help="N>1: put the cond/uncond forwards on two Ulysses groups of N/2 ranks even when "
                         "the head count would allow one group of N (what N=8 does for 12 heads)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="replay the DiT forward as a CUDA graph; default: on when the forward is "
                         "sequence-parallel (kernels a fraction of a millisecond long), off otherwise")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
