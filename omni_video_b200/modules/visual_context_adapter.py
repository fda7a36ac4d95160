"""B200-native VisualContextAdapter (drop-in for reference omnivideo/modules/visual_context_adapter.py).

VAE latents of the conditioning clip -> context tokens for the DiT's cross-attention:

    reference (visual_context_adapter.py:77-115)                      kernel
    ----------------------------------------------------------------  --------------------------
    patch_embedding: Conv3d(C_in, hidden, k=s=patch)  [bf16 autocast]  ovb_patchify3d + ovb_gemm_bf16
    norm1: LayerNorm(hidden), no affine               [fp32]           ovb_layernorm (scale=shift=0)
    projection: Linear(hidden, out_dim)               [bf16 autocast]  ovb_gemm_bf16
    norm2: LayerNorm(out_dim), no affine              [fp32]           ovb_layernorm -> fp32

Parameter names/shapes are the reference's (`patch_embedding.{weight,bias}`, `projection.{weight,bias}`),
so `visual_context_adapter_pytorch_model.bin` (unified_model.py:282) loads unchanged.  Rounding points
follow the reference's bf16-autocast run: both GEMM outputs are bf16, both LayerNorms take fp32
statistics; the result is returned in fp32 like `nn.LayerNorm` under autocast.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops

bf16 = torch.bfloat16
f32 = torch.float32


class VisualContextAdapter(nn.Module):
    def __init__(self, patch_size=(1, 4, 4), in_channels=16, hidden_dim=2048, out_dim=4098, eps=1e-6):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.in_channels, self.hidden_dim, self.out_dim, self.eps = in_channels, hidden_dim, out_dim, eps
        self.patch_embedding = nn.Conv3d(in_channels, hidden_dim, kernel_size=self.patch_size,
                                         stride=self.patch_size)
        self.norm1 = nn.LayerNorm(hidden_dim, eps=eps, elementwise_affine=False)
        self.projection = nn.Linear(hidden_dim, out_dim)
        self.norm2 = nn.LayerNorm(out_dim, eps=eps, elementwise_affine=False)
        self.apply(self._init_weights)

    def _init_weights(self, module):
        """reference visual_context_adapter.py:64-75"""
        if isinstance(module, nn.Conv3d):
            nn.init.xavier_uniform_(module.weight.flatten(1))
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.Linear):
            nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                nn.init.zeros_(module.bias)

    def forward(self, x):
        """x: list of [C_in, F, H, W] | [B, C_in, F, H, W] | [C_in, F, H, W]  ->  fp32 [B, N, out_dim]."""
        return visual_context_forward(self, x)


def _weights(mod):
    """Packed bf16 copies of the four parameters: a derived cache kept on the module, rebuilt when
    the parameters move or change."""
    ps = (mod.patch_embedding.weight, mod.patch_embedding.bias, mod.projection.weight, mod.projection.bias)
    key = tuple((p.data_ptr(), p._version, p.device, p.dtype) for p in ps)
    packed = getattr(mod, "_ovb_packed", None)
    if packed is None or packed[0] != key:
        packed = (key, ps[0].detach().flatten(1).to(bf16).contiguous(), ps[1].detach().to(bf16).contiguous(),
                  ps[2].detach().to(bf16).contiguous(), ps[3].detach().to(bf16).contiguous())
        mod._ovb_packed = packed
    return packed[1:]


@torch.no_grad()
def visual_context_forward(mod, x):
    """`mod`: our VisualContextAdapter or the reference's own instance (same parameter tree) —
    omni_video_b200.patch.patch_reference_unified binds this as its `forward`."""
    if isinstance(x, list):
        x = torch.stack(list(x))
    elif x.dim() == 4:
        x = x.unsqueeze(0)
    if x.device.type != "cuda":
        raise RuntimeError("VisualContextAdapter (B200 path) needs CUDA tensors; there is no CPU fallback")
    w0, b0, w1, b1 = _weights(mod)
    if x.dtype not in (f32, bf16):
        x = x.float()
    patch = tuple(mod.patch_size)
    toks = torch.cat([ops.patchify3d(u.contiguous(), patch) for u in x], dim=0)
    n = toks.shape[0] // x.shape[0]
    h = ops.gemm(toks, w0, b0)                                               # bf16 [B*N, hidden]
    zero_h = torch.zeros((1, mod.hidden_dim), device=x.device, dtype=f32)
    h = ops.layernorm_modulate(h, zero_h, zero_h, mod.eps)                   # fp32 stats -> bf16
    h = ops.gemm(h, w1, b1)                                                  # bf16 [B*N, out_dim]
    zero_o = torch.zeros((1, mod.out_dim), device=x.device, dtype=f32)
    out = ops.layernorm_modulate(h, zero_o, zero_o, mod.eps, out_dtype=f32)
    return out.view(x.shape[0], n, mod.out_dim)
