"""The denoising loop of the reference pipeline over the B200 modules.

Reference: `OmniVideoX2XUnified.generate`, omnivideo/x2x_gen_unified.py:403-492 —

    for t in timesteps:
        model = high-noise expert if t >= boundary * 1000 else low-noise expert        (:228-258, :477)
        cond   = model(latents, t, **arg_c)[0] ;  uncond = model(latents, t, **arg_null)[0]   (:480-481)
        noise_pred = uncond + guide_scale * (cond - uncond)                            (:484)
        latents = scheduler.step(noise_pred, t, latents)                                (:486-491)

What changes here is only HOW the same arithmetic is scheduled:

* **CFG batching** — cond and uncond share latents, timestep and weights, so they run as ONE forward with
  B = 2 (`WanModel.forward` takes lists): half the launches and weight reads, twice the CTAs per attention
  launch.  With `cfg_pair=(branch, group)` the two branches instead run on two process groups (each possibly
  a sequence-parallel group) and exchange one latent-sized tensor per step.
* **Expert residency** — the reference moves the expert it does not need to the CPU at the boundary
  (`_prepare_model_for_timestep`, offload_model / init_on_cpu); 2 x 28 GB of bf16 weights fit a 180 GB B200
  next to the activations, so both stay resident and the switch is a Python reference.
* **Fused update** — guidance, x0 prediction, UniC corrector and UniP predictor are single-launch linear
  combinations (`FlowUniPCMultistepScheduler.step_cfg`); likewise the DPM-Solver++ update of
  `sample_solver='dpm++'` (`FlowDPMSolverMultistepScheduler.step_cfg`).

The conditioning front-end consumes `random.random()` exactly as the reference does whether the two branches
are batched or not (2 draws per sample per call: 4 per step either way).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch

from .utils.fm_solvers import FlowDPMSolverMultistepScheduler, get_sampling_sigmas, retrieve_timesteps
from .utils.fm_solvers_unipc import FlowUniPCMultistepScheduler

__all__ = ["DenoiseLoop", "merge_cfg_args"]

# keyword arguments of UnifiedWanWithMixedConditionModel.forward that hold one entry per sample, and the
# rank of ONE sample's tensor (a tensor of rank + 1 carries a leading batch dimension, unified_model.py:226-233)
_PER_SAMPLE = {"context": 2, "ar_vision_input": 2, "visual_emb": 4}


def _as_list(v, sample_rank):
    if v is None or isinstance(v, list):
        return v
    if isinstance(v, torch.Tensor):
        return [v[i] for i in range(v.shape[0])] if v.dim() == sample_rank + 1 else [v]
    return list(v)


def merge_cfg_args(arg_c: Dict, arg_null: Dict) -> Optional[Dict]:
    """One keyword set for a B = 2 call [cond sample, uncond sample], or None when the two calls differ in
    something that is not per-sample (then they cannot share a forward)."""
    merged = {}
    for k in set(arg_c) | set(arg_null):
        a, b = arg_c.get(k), arg_null.get(k)
        if k in _PER_SAMPLE:
            la, lb = _as_list(a, _PER_SAMPLE[k]), _as_list(b, _PER_SAMPLE[k])
            if la is None and lb is None:
                merged[k] = None
            elif la is None or lb is None or len(la) != 1 or len(lb) != 1:
                return None
            else:
                merged[k] = [la[0], lb[0]]
        elif a is b or (not isinstance(a, (torch.Tensor, dict, list)) and a == b):
            merged[k] = a
        else:
            return None
    return merged


class DenoiseLoop:
    """`high_noise_model` / `low_noise_model`: callables with the reference's convention
    `model(list_of_latents, t=timestep[B], **kwargs) -> list of fp32 [C, F, H, W]` (our
    `UnifiedWanWithMixedConditionModel`, a bare `WanModel`, or the reference's own modules after
    `patch.patch_reference_unified`).  One expert: pass the same object twice or `low_noise_model=None`."""

    def __init__(self, high_noise_model, low_noise_model=None, boundary: float = 0.875,
                 num_train_timesteps: int = 1000, batch_cfg: bool = True,
                 cfg_pair: Optional[Tuple[int, object]] = None):
        self.high_noise_model = high_noise_model
        self.low_noise_model = low_noise_model if low_noise_model is not None else high_noise_model
        self.boundary = boundary
        self.num_train_timesteps = num_train_timesteps
        self.batch_cfg = batch_cfg and cfg_pair is None
        self.cfg_pair = cfg_pair
        self.calls = {"high_noise_model": 0, "low_noise_model": 0}     # forwards issued per expert

    def select_model(self, t) -> Tuple[str, Callable]:
        """x2x_gen_unified.py:236-241 without the offload: both experts stay on the device."""
        name = "high_noise_model" if float(t) >= self.boundary * self.num_train_timesteps else "low_noise_model"
        return name, getattr(self, name)

    def predict(self, latent: torch.Tensor, timestep: torch.Tensor, arg_c: Dict, arg_null: Dict):
        """(cond, uncond) noise predictions for one latent at one timestep."""
        name, model = self.select_model(timestep[0])
        if self.cfg_pair is not None:
            import torch.distributed as dist
            branch, group = self.cfg_pair
            mine = model([latent], t=timestep, **(arg_c if branch == 0 else arg_null))[0]
            self.calls[name] += 1
            both = torch.empty((2,) + tuple(mine.shape), device=mine.device, dtype=mine.dtype)
            dist.all_gather_into_tensor(both, mine.contiguous(), group=group)
            return both[0], both[1]
        merged = merge_cfg_args(arg_c, arg_null) if self.batch_cfg else None
        if merged is not None:
            cond, uncond = model([latent, latent], t=torch.cat([timestep, timestep]), **merged)
            self.calls[name] += 1
            return cond, uncond
        cond = model([latent], t=timestep, **arg_c)[0]
        uncond = model([latent], t=timestep, **arg_null)[0]
        self.calls[name] += 2
        return cond, uncond

    @torch.no_grad()
    def sample(self, noise: torch.Tensor, arg_c: Dict, arg_null: Dict, *, sampling_steps: int = 50,
               shift: float = 5.0, guide_scale: float = 5.0, sample_solver: str = "unipc",
               callback: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None) -> List[torch.Tensor]:
        """noise: [C, F, H, W] on the device (x2x_gen_unified.py:382-391) -> [latents], ready for
        `Wan2_1_VAE.decode`.  arg_c / arg_null: the keyword sets of :452-472 (for a bare WanModel:
        `context` and `seq_len`)."""
        if sample_solver == "unipc":                                   # x2x_gen_unified.py:409-416
            sched = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1,
                                                use_dynamic_shifting=False)
            sched.set_timesteps(sampling_steps, device=noise.device, shift=shift)
            timesteps = sched.timesteps
        elif sample_solver == "dpm++":                                 # :417-427
            sched = FlowDPMSolverMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1,
                                                    use_dynamic_shifting=False)
            timesteps, _ = retrieve_timesteps(sched, device=noise.device,
                                              sigmas=get_sampling_sigmas(sampling_steps, shift))
        else:
            raise NotImplementedError("Unsupported solver.")
        latent = noise
        for i, t in enumerate(timesteps):
            timestep = t.unsqueeze(0)
            cond, uncond = self.predict(latent, timestep, arg_c, arg_null)
            latent = sched.step_cfg(cond, uncond, guide_scale, t, latent.unsqueeze(0))[0].squeeze(0)
            if callback is not None:
                callback(i, t, latent)
        return [latent]
