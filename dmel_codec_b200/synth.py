"""Seeded synthetic utterances.

Every utterance is a function of its id alone (``seed = 1234 + utterance_id``),
so any rank regenerates its own shard and nothing moves between GPUs.  Values
lie in [-0.95, 0.95], mirroring the reference data module's
``librosa.util.normalize(audio) * 0.95`` (reference
dataset/lhotse_tts_dataset.py:32).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

BASE_SEED = 1234


def _colour(noise: torch.Tensor, pole: float = 0.95, taps: int = 96) -> torch.Tensor:
    """Leaky integrator as a truncated FIR (pole**k), so it vectorises."""
    k = torch.arange(taps, dtype=noise.dtype, device=noise.device)
    kernel = torch.pow(torch.tensor(pole, dtype=noise.dtype, device=noise.device), k).flip(0)
    x = torch.nn.functional.pad(noise[None, None, :], (taps - 1, 0))
    return torch.nn.functional.conv1d(x, kernel[None, None, :])[0, 0]


def utterance(utt_id: int, n_samples: int, sample_rate: int, kind: str = "speech",
              device: str | torch.device = "cpu") -> torch.Tensor:
    """One fp32 waveform of ``n_samples``.

    kind="speech": coloured Gaussian noise under a slow sin^2 envelope,
    peak-normalised to 0.95, with one span of exact digital silence (drives the
    log clamp floor).  kind="noise": 0.1 * randn clipped to +-0.95 (densest
    bin-edge traffic).  Generated on the CPU generator so the same id gives the
    same samples on every machine; ``device`` only says where the result lands.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(BASE_SEED + int(utt_id))
    white = torch.randn(n_samples, generator=g, dtype=torch.float32)
    if kind == "noise":
        return (0.1 * white).clamp_(-0.95, 0.95).to(device)
    if kind != "speech":
        raise ValueError(f"unknown synthetic kind {kind!r}")
    x = _colour(white)
    t = torch.arange(n_samples, dtype=torch.float32) / float(sample_rate)
    rate = 0.7 + 0.6 * torch.rand(1, generator=g).item()  # syllable-ish envelope, Hz
    phase = 2 * math.pi * torch.rand(1, generator=g).item()
    x = x * torch.sin(math.pi * rate * t + phase).pow(2)
    x = x + 0.02 * white  # broadband floor so high mel channels are not empty
    x = x * (0.95 / x.abs().max().clamp_min(1e-12))
    # one exact-zero span: 5 % of the utterance, placed by the seed
    span = max(1, n_samples // 20)
    start = int(torch.randint(0, max(1, n_samples - span), (1,), generator=g).item())
    x[start:start + span] = 0.0
    return x.to(device)


def batch(utt_ids: Sequence[int], n_samples: int, sample_rate: int, kind: str = "speech",
          lengths: Optional[Sequence[int]] = None, device: str | torch.device = "cpu",
          pin: bool = False) -> torch.Tensor:
    """(B, 1, n_samples) fp32, right zero-padded past ``lengths[i]`` the way the
    reference collate does (reference dataset/lhotse_tts_dataset.py:46-65)."""
    out = torch.zeros(len(utt_ids), 1, n_samples, dtype=torch.float32)
    for row, uid in enumerate(utt_ids):
        n = n_samples if lengths is None else int(lengths[row])
        out[row, 0, :n] = utterance(uid, n, sample_rate, kind)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    return out.to(device) if str(device) != "cpu" else out


def device_batch(utt_ids: Sequence[int], n_samples: int, sample_rate: int,
                 device: str | torch.device) -> torch.Tensor:
    """Large-scale variant for the dataset-size benches: same recipe evaluated
    with the device generator (one seed per utterance, different stream from the
    CPU generator, so NOT sample-identical to ``batch``).  (B, 1, n_samples)."""
    dev = torch.device(device)
    out = torch.empty(len(utt_ids), 1, n_samples, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    t = torch.arange(n_samples, dtype=torch.float32, device=dev) / float(sample_rate)
    span = max(1, n_samples // 20)
    for row, uid in enumerate(utt_ids):
        g.manual_seed(BASE_SEED + int(uid))
        white = torch.randn(n_samples, generator=g, dtype=torch.float32, device=dev)
        x = _colour(white)
        rate = 0.7 + 0.6 * ((uid * 2654435761) % 1000) / 1000.0
        phase = 2 * math.pi * ((uid * 40503) % 1000) / 1000.0
        x = x * torch.sin(math.pi * rate * t + phase).pow(2) + 0.02 * white
        x = x * (0.95 / x.abs().max().clamp_min(1e-12))
        start = (uid * 7919) % max(1, n_samples - span)
        x[start:start + span] = 0.0
        out[row, 0] = x
    return out
