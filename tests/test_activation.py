"""BigVGAN's anti-aliased Snake activation: the oracle against golden vectors produced by the
reference's own modules, the CUDA kernel against both."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import activation_oracle as A

CASES = ["snakebeta_small", "snakebeta_long", "snake_shared_param", "shorter_than_filter", "single_sample", "large_arguments"]
TOL = 2e-5  # |a - b| <= TOL * max(1, |b|): float32 FIR sums of 12 terms twice, and sin^2 through a reduced argument


def gate(log_alpha, log_beta):
    """The Snake amplifies a rounding difference du of the upsampled value by |d/du sin^2(a u) / b| <= a / b: two float32
    evaluations of the same FIR sum (another summation order) differ by an ulp, so the gate scales with max(a / b) once
    that exceeds 1.  (The reference's own fused kernel, built with --use_fast_math, is looser still.)"""
    amp = (torch.exp(torch.as_tensor(log_alpha).double()) / torch.exp(torch.as_tensor(log_beta).double())).max().item()
    return TOL * max(1.0, amp)


@pytest.fixture(scope="module")
def act_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "activation_golden.npz"))


def _close(a, b, tol):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs() / torch.clamp(b.abs(), min=1.0)).max().item() <= tol


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_modules(act_golden, name):
    g = {k: torch.from_numpy(act_golden[f"{name}/{k}"]) for k in ("x", "log_alpha", "log_beta", "up_taps", "down_taps", "y")}
    assert torch.equal(A.kaiser_sinc_filter(), g["up_taps"]) and torch.equal(A.kaiser_sinc_filter(), g["down_taps"])
    y = A.anti_alias_snake(g["x"], g["log_alpha"], g["log_beta"])
    assert y.shape == g["y"].shape
    assert _close(y, g["y"], 2e-6), ((y - g["y"]).abs().max().item())


def test_filter_is_a_unit_gain_low_pass():
    taps = A.kaiser_sinc_filter()
    assert taps.numel() == 12 and abs(taps.sum().item() - 1.0) < 1e-6 and torch.allclose(taps, taps.flip(0))
    # a constant passes both resamplers unchanged (DC gain 1, replicate padding), so the module maps c -> snake(c)
    x = torch.full((1, 2, 40), 0.75)
    la, lb = torch.tensor([0.3, -0.2]), torch.tensor([0.1, 0.4])
    want = A.snake_beta(x, la, lb)
    assert torch.allclose(A.anti_alias_snake(x, la, lb), want, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_activation_matches_reference_golden(native_lib, act_golden, name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dmel_codec_b200 as d
    g = {k: torch.from_numpy(act_golden[f"{name}/{k}"]) for k in ("x", "log_alpha", "log_beta", "y")}
    tie = name.startswith("snake_")
    mod = d.AntiAliasSnake(g["x"].shape[1], tie_beta=tie).cuda()
    mod.alpha.copy_(g["log_alpha"])
    if not tie:
        mod.beta.copy_(g["log_beta"])
    y = mod(g["x"].cuda()).cpu()
    assert _close(y, g["y"], gate(g["log_alpha"], g["log_beta"])), ((y - g["y"]).abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 5, 1008), (1, 3, 1009), (1, 2, 1010), (1, 2, 1011), (3, 4, 2016 + 7), (1, 1, 4099), (2, 2, 3),
                                   (1, 2, 1), (2, 3, 236), (1, 3, 239), (2, 2, 240), (1, 2, 241), (1, 2, 242), (1, 2, 243), (1, 1, 244),
                                   (2, 3, 480), (1, 2, 481), (1, 2, 482), (3, 7, 724), (1, 1, 100000), (8, 64, 5000), (3, 50, 20001)])
def test_cuda_activation_matches_oracle_across_tile_boundaries(native_lib, shape):
    """row lengths around the 240-output warp tile (and the 1008 of the first version): the downsampler's replicate
    padding falls into the last tile, the one before it (last tile of one or two outputs), or both; lengths that are
    not multiples of four take the unaligned load / store path; several rows and channels walk the persistent grid's
    tile -> (row, channel) arithmetic"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dmel_codec_b200 as d
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen) * 1.2
    la, lb = torch.randn(shape[1], generator=gen) * 0.8, torch.randn(shape[1], generator=gen) * 0.5
    mod = d.AntiAliasSnake(shape[1]).cuda()
    mod.alpha.copy_(la)
    mod.beta.copy_(lb)
    y = mod(x.cuda()).cpu()
    want = A.anti_alias_snake(x, la, lb)
    assert _close(y, want, gate(la, lb)), ((y - want).abs().max().item())
    # a constant row passes the resamplers unchanged
    c = torch.full(shape, -0.4)
    assert _close(mod(c.cuda()).cpu(), A.snake_beta(c, la, lb), gate(la, lb))
