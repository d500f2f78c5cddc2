"""BigVGAN's anti-aliased Snake activation in one kernel.

Drop-in for ``Activation1d(activation=SnakeBeta(channels, alpha_logscale=True))`` of the reference
(models/modules/bigvgan/alias_free_activation/torch/act.py:8-29 and .../cuda/activation1d.py:35-77, activations.py:56-111):
same constructor idea, same ``forward(x)`` on (B, C, T), inference only like the reference's fused path
(activation1d.py:29-32 raises in backward).  The reference's own fused kernel is compiled for sm_70 / sm_80 without PTX
(.../cuda/load.py:21-24) and cannot run on a B200; its torch path makes three HBM round trips over a 2x-sized tensor.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _native


def kaiser_sinc_taps(cutoff: float = 0.25, half_width: float = 0.3, kernel_size: int = 12) -> np.ndarray:
    """The low-pass taps both resamplers of the reference use (.../torch/filter.py:31-60), float32."""
    half_size = kernel_size // 2
    a = 2.285 * (half_size - 1) * math.pi * 4 * half_width + 7.95
    beta = 0.1102 * (a - 8.7) if a > 50.0 else (0.5842 * (a - 21) ** 0.4 + 0.07886 * (a - 21.0) if a >= 21.0 else 0.0)
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = torch.arange(-half_size, half_size) + 0.5 if kernel_size % 2 == 0 else torch.arange(kernel_size) - half_size
    taps = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (taps / taps.sum()).to(torch.float32).numpy()


class AntiAliasSnake(nn.Module):
    """up 2x -> ``x + sin^2(x exp(alpha)) / (exp(beta) + 1e-9)`` -> down 2x, per channel.

    ``alpha`` / ``beta`` are log-scale parameters as in ``SnakeBeta(alpha_logscale=True)``; ``tie_beta=True`` gives the
    one-parameter ``Snake``.  ``from_reference(act1d)`` copies filters and parameters from a reference ``Activation1d``."""

    def __init__(self, channels: int, tie_beta: bool = False):
        super().__init__()
        self.channels = int(channels)
        self.alpha = nn.Parameter(torch.zeros(channels), requires_grad=False)
        self.beta = None if tie_beta else nn.Parameter(torch.zeros(channels), requires_grad=False)
        self.up_taps = np.ascontiguousarray(kaiser_sinc_taps(), dtype=np.float32)
        self.down_taps = np.ascontiguousarray(kaiser_sinc_taps(), dtype=np.float32)

    @classmethod
    def from_reference(cls, act1d) -> "AntiAliasSnake":
        act = act1d.act
        tie = not hasattr(act, "beta")
        mod = cls(act.alpha.numel(), tie_beta=tie)
        log = (lambda p: p.detach().float()) if act.alpha_logscale else (lambda p: torch.log(p.detach().float()))
        mod.alpha.copy_(log(act.alpha))
        if not tie:
            mod.beta.copy_(log(act.beta))
        mod.up_taps = np.ascontiguousarray(act1d.upsample.filter.reshape(-1).float().cpu().numpy())
        mod.down_taps = np.ascontiguousarray(act1d.downsample.lowpass.filter.reshape(-1).float().cpu().numpy())
        if mod.up_taps.size != 12 or mod.down_taps.size != 12:
            raise NotImplementedError("the kernel is built for the reference's 12-tap filters (ratio 2)")
        return mod.to(act.alpha.device)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("AntiAliasSnake runs on CUDA tensors only (no CPU fallback)")
        if x.ndim != 3 or x.shape[1] != self.channels:
            raise ValueError(f"expected (B, {self.channels}, T), got {tuple(x.shape)}")
        xs = x.float().contiguous()
        alpha = self.alpha.to(device=xs.device, dtype=torch.float32).contiguous()
        beta = alpha if self.beta is None else self.beta.to(device=xs.device, dtype=torch.float32).contiguous()
        y = torch.empty_like(xs) if out is None else out
        if y.shape != xs.shape or y.dtype != torch.float32 or not y.is_contiguous() or y.device != xs.device:
            raise ValueError("out must be a contiguous float32 tensor of the input's shape on the input's device")
        b, c, t = xs.shape
        if xs.numel():
            _native.check(_native.load().dmel_antialias_snake_f32(
                xs.data_ptr(), b, c, t, self.up_taps.ctypes.data, self.down_taps.ctypes.data, alpha.data_ptr(), beta.data_ptr(),
                y.data_ptr(), torch.cuda.current_stream(xs.device).cuda_stream))
        return y
