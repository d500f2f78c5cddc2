"""CPU oracle of BigVGAN's anti-aliased Snake activation (second half).  TEST INFRASTRUCTURE ONLY.

Restates, with plain torch CPU ops, the reference's torch path

    Activation1d.forward            models/modules/bigvgan/alias_free_activation/torch/act.py:24-29
      UpSample1d(ratio 2, K 12)     .../torch/resample.py:10-36   replicate pad 5, 2 * conv_transpose1d(stride 2), crop 15 | 15
      SnakeBeta / Snake             models/modules/bigvgan/activations.py:45-53, :101-111   x + sin^2(a x) / (b + 1e-9)
      DownSample1d(ratio 2, K 12)   .../torch/resample.py:39-58, filter.py:63-101   replicate pad 5 | 6, conv1d(stride 2)
      kaiser_sinc_filter1d          .../torch/filter.py:31-60     cutoff 0.25, half width 0.3, 12 taps, sum 1

which is also what the reference's fused sm_70/sm_80 CUDA kernel computes
(.../cuda/anti_alias_activation_cuda.cu:44-179, hyper-parameters hard-coded, alpha / beta given in log scale).
Pinned: tests/golden/activation_golden.npz is produced by tests/golden/make_activation_golden.py from the
reference's own modules, and tests/test_activation.py holds this oracle to it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

KERNEL = 12


def kaiser_sinc_filter(cutoff: float = 0.25, half_width: float = 0.3, kernel_size: int = KERNEL) -> torch.Tensor:
    """(kernel_size,) float32 low-pass taps (filter.py:31-60): Kaiser-windowed sinc, normalised to sum 1."""
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    a = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if a > 50.0:
        beta = 0.1102 * (a - 8.7)
    elif a >= 21.0:
        beta = 0.5842 * (a - 21) ** 0.4 + 0.07886 * (a - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        time = torch.arange(-half_size, half_size) + 0.5
    else:
        time = torch.arange(kernel_size) - half_size
    taps = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (taps / taps.sum()).to(torch.float32)


def upsample2(x: torch.Tensor, taps: torch.Tensor) -> torch.Tensor:
    """(B, C, T) -> (B, C, 2T) (resample.py:28-36)"""
    c = x.shape[1]
    k = taps.numel()
    pad = k // 2 - 1
    pad_left = pad * 2 + (k - 2) // 2
    pad_right = pad * 2 + (k - 2 + 1) // 2
    y = F.pad(x, (pad, pad), mode="replicate")
    y = 2 * F.conv_transpose1d(y, taps.view(1, 1, k).expand(c, -1, -1), stride=2, groups=c)
    return y[..., pad_left:-pad_right]


def downsample2(x: torch.Tensor, taps: torch.Tensor) -> torch.Tensor:
    """(B, C, 2T) -> (B, C, T) (filter.py:93-101 with stride 2)"""
    c = x.shape[1]
    k = taps.numel()
    y = F.pad(x, (k // 2 - 1, k // 2), mode="replicate")
    return F.conv1d(y, taps.view(1, 1, k).expand(c, -1, -1), stride=2, groups=c)


def snake_beta(x: torch.Tensor, log_alpha: torch.Tensor, log_beta: torch.Tensor) -> torch.Tensor:
    """x + sin^2(x * exp(alpha)) / (exp(beta) + 1e-9), per channel (activations.py:101-111 with alpha_logscale=True;
    Snake is the case beta = alpha)."""
    a = torch.exp(log_alpha)[None, :, None]
    b = torch.exp(log_beta)[None, :, None]
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a).pow(2)


def anti_alias_snake(x: torch.Tensor, log_alpha: torch.Tensor, log_beta: torch.Tensor, up_taps=None, down_taps=None) -> torch.Tensor:
    """Activation1d.forward (act.py:24-29) with a SnakeBeta in log scale: up 2x, activate, down 2x."""
    up_taps = kaiser_sinc_filter() if up_taps is None else up_taps
    down_taps = kaiser_sinc_filter() if down_taps is None else down_taps
    return downsample2(snake_beta(upsample2(x.float(), up_taps), log_alpha, log_beta), down_taps)
