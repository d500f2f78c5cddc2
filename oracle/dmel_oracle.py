"""CPU oracle for the dMel tokenization hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  Nothing under
``dmel_codec_b200/`` imports it; the product path is CUDA-only and fails loudly
without its extension.

Parity status
-------------
* waveform -> log-mel: PINNED.  ``tests/golden/make_golden.py`` executes the
  reference's own ``dmel_codec/utils/spectrogram.py`` (unmodified, loaded by
  path in the build container) and commits its outputs under ``tests/golden/``;
  ``tests/test_oracle.py`` checks this restatement against those fixtures.
  One caveat: the reference gets its filterbank from ``librosa.filters.mel``
  (librosa==0.10.2.post1, reference ``setup.py:15``), which is not installed
  here and cannot be.  ``slaney_filterbank`` below restates that published
  algorithm; it is cross-checked against
  ``torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")``.
* dMel bin quantizer / dequantizer / calibration: PARITY UNPINNED.  The
  reference contains no such code; the spec is SURVEY.md
  Appendix B and this file is its executable statement.

All functions are plain torch/numpy on CPU tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

LOG_FLOOR_CLIP = 1e-5  # reference utils/spectrogram.py:38 (clip_val)
MAG_EPS = 1e-9  # reference utils/spectrogram.py:76


# --------------------------------------------------------------------------- #
# filterbank + window  (reference utils/spectrogram.py:43-56)
# --------------------------------------------------------------------------- #
def _hz_to_slaney_mel(hz: np.ndarray) -> np.ndarray:
    """Slaney (Auditory Toolbox) mel scale, the ``htk=False`` default of
    ``librosa.filters.mel`` used at reference utils/spectrogram.py:45-51."""
    hz = np.asarray(hz, dtype=np.float64)
    lin_step = 200.0 / 3.0
    knee_hz = 1000.0
    knee_mel = knee_hz / lin_step
    log_step = math.log(6.4) / 27.0
    out = hz / lin_step
    above = hz >= knee_hz
    safe = np.where(above, hz, knee_hz)
    return np.where(above, knee_mel + np.log(safe / knee_hz) / log_step, out)


def _slaney_mel_to_hz(mel: np.ndarray) -> np.ndarray:
    mel = np.asarray(mel, dtype=np.float64)
    lin_step = 200.0 / 3.0
    knee_hz = 1000.0
    knee_mel = knee_hz / lin_step
    log_step = math.log(6.4) / 27.0
    return np.where(mel >= knee_mel, knee_hz * np.exp(log_step * (mel - knee_mel)), lin_step * mel)


def slaney_filterbank(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0,
                      f_max: Optional[float] = None) -> np.ndarray:
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with its defaults
    (slaney scale, slaney area-normalisation, float32 result) as called at
    reference utils/spectrogram.py:45-51.  float64 math, float32 storage of the
    triangles *before* the normalisation multiply."""
    if f_max is None:
        f_max = float(sample_rate) / 2.0
    n_freq = 1 + n_fft // 2
    bin_hz = np.fft.rfftfreq(n_fft, d=1.0 / sample_rate)
    edges_hz = _slaney_mel_to_hz(
        np.linspace(_hz_to_slaney_mel(np.float64(f_min)), _hz_to_slaney_mel(np.float64(f_max)), n_mels + 2))
    widths = np.diff(edges_hz)
    dist = edges_hz[:, None] - bin_hz[None, :]
    bank = np.zeros((n_mels, n_freq), dtype=np.float32)
    for m in range(n_mels):
        rising = -dist[m] / widths[m]
        falling = dist[m + 2] / widths[m + 1]
        bank[m] = np.maximum(0.0, np.minimum(rising, falling))
    area = 2.0 / (edges_hz[2:n_mels + 2] - edges_hz[:n_mels])
    bank *= area[:, None]
    return bank


def stft_window(win_length: int, n_fft: int) -> torch.Tensor:
    """Periodic Hann of ``win_length`` (reference utils/spectrogram.py:53),
    centred inside ``n_fft`` the way ``torch.stft`` does when shorter."""
    w = torch.hann_window(win_length)
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - win_length - left))
    return w


# --------------------------------------------------------------------------- #
# waveform -> log-mel  (reference utils/spectrogram.py:41-81)
# --------------------------------------------------------------------------- #
@dataclass(frozen=True)
class MelConfig:
    sample_rate: int = 44100
    n_fft: int = 2048
    win_length: int = 2048
    hop_length: int = 512
    n_mels: int = 128
    f_min: float = 0.0
    f_max: Optional[float] = None
    center: bool = False

    @property
    def pad(self) -> int:
        return (self.n_fft - self.hop_length) // 2  # utils/spectrogram.py:58

    def n_frames(self, n_samples: int) -> int:
        padded = n_samples + 2 * self.pad + (2 * (self.n_fft // 2) if self.center else 0)
        return 1 + (padded - self.n_fft) // self.hop_length


def reflect_index(i: np.ndarray, n: int) -> np.ndarray:
    """Index map of ``F.pad(mode='reflect')`` (no edge repeat):
    i<0 -> -i, i>=n -> 2(n-1)-i.  Valid for -n < i < 2n-1."""
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def magnitude_frames(wav: torch.Tensor, cfg: MelConfig) -> torch.Tensor:
    """(B, L) fp32 -> (B, F, T) fp32 magnitudes: reflect pad (:58-62), frame,
    window, real FFT (:64-75), sqrt(re^2 + im^2 + 1e-9) (:76)."""
    if wav.ndim == 3:
        wav = wav.squeeze(1)
    wav = wav.to(torch.float32)
    n = wav.shape[-1]
    if cfg.pad >= n:
        raise ValueError(f"reflect pad {cfg.pad} needs more than {cfg.pad} samples, got {n}")
    padded = torch.nn.functional.pad(wav[:, None, :], (cfg.pad, cfg.pad), mode="reflect")[:, 0, :]
    if cfg.center:
        h = cfg.n_fft // 2
        padded = torch.nn.functional.pad(padded[:, None, :], (h, h), mode="reflect")[:, 0, :]
    frames = padded.unfold(-1, cfg.n_fft, cfg.hop_length)  # (B, T, n_fft)
    spec = torch.fft.rfft(frames * stft_window(cfg.win_length, cfg.n_fft), dim=-1)  # (B, T, F)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + MAG_EPS)
    return mag.transpose(1, 2).contiguous()


def log_mel(wav: torch.Tensor, cfg: MelConfig, bank: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, L) or (B, 1, L) fp32 -> (B, n_mels, T) fp32 natural-log mel
    (reference utils/spectrogram.py:41-81; compression at :38-39)."""
    if bank is None:
        # the reference hands the *raw* f_max down (utils/spectrogram.py:114), so None -> sr/2
        bank = torch.from_numpy(slaney_filterbank(cfg.sample_rate, cfg.n_fft, cfg.n_mels, cfg.f_min, cfg.f_max))
    mel = torch.matmul(bank, magnitude_frames(wav, cfg))  # :78
    return torch.log(torch.clamp(mel, min=LOG_FLOOR_CLIP))  # :38-39, C = 1


def valid_frames(lengths: torch.Tensor, hop_length: int) -> torch.Tensor:
    """Caller's rule for how many frames of each row are real audio:
    ``mel_lengths = audio_lengths // hop_length`` (reference
    models/codec_lit_modules.py:176, mask at utils/utils.py:48-55)."""
    return torch.div(lengths, hop_length, rounding_mode="floor")


# --------------------------------------------------------------------------- #
# dMel quantizer  (NOT in the reference; SURVEY.md Appendix B is the spec)
# --------------------------------------------------------------------------- #
def calibrate_minmax(mel: torch.Tensor, n_valid: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel min / max of (B, M, T) log-mel over valid frames
    (t < n_valid[b]).  Exact, order independent."""
    b, m, t = mel.shape
    if n_valid is None:
        return mel.amin(dim=(0, 2)), mel.amax(dim=(0, 2))
    keep = (torch.arange(t)[None, :] < n_valid.reshape(b, 1))[:, None, :].expand(b, m, t)
    lo = torch.where(keep, mel, torch.full_like(mel, float("inf"))).amin(dim=(0, 2))
    hi = torch.where(keep, mel, torch.full_like(mel, float("-inf"))).amax(dim=(0, 2))
    return lo, hi


def bin_scale(lo: torch.Tensor, hi: torch.Tensor, n_bins: int) -> torch.Tensor:
    """s_c = K / (hi_c - lo_c) in fp32, 0 for a degenerate channel."""
    width = hi - lo
    return torch.where(width > 0, torch.tensor(float(n_bins), dtype=torch.float32) / width,
                       torch.zeros_like(width))


def dmel_encode(mel: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, n_bins: int) -> torch.Tensor:
    """code = clamp(floor((x - lo_c) * s_c), 0, K-1) as uint8, fp32, this op order."""
    s = bin_scale(lo, hi, n_bins)
    q = torch.floor((mel - lo[None, :, None]) * s[None, :, None])
    return torch.clamp(q, 0, n_bins - 1).to(torch.uint8)


def dmel_decode_table(lo: torch.Tensor, hi: torch.Tensor, n_bins: int) -> torch.Tensor:
    """(M, K) bin centres: lo_c + (k + 0.5) * ((hi_c - lo_c) / K)."""
    step = (hi - lo) / float(n_bins)
    k = torch.arange(n_bins, dtype=torch.float32) + 0.5
    return lo[:, None] + k[None, :] * step[:, None]


def dmel_decode(codes: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, n_bins: int) -> torch.Tensor:
    table = dmel_decode_table(lo, hi, n_bins)  # (M, K)
    idx = codes.long()
    m = codes.shape[1]
    return table[torch.arange(m)[None, :, None], idx]


def interior_edge_distance(mel: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, n_bins: int) -> torch.Tensor:
    """Distance (in log-mel units) from each value to the nearest INTERIOR bin
    edge lo_c + i*step_c, i = 1..K-1 (fp64; used to classify code mismatches)."""
    x = mel.double()
    lo64, hi64 = lo.double()[None, :, None], hi.double()[None, :, None]
    step = (hi64 - lo64) / n_bins
    pos = (x - lo64) / torch.where(step > 0, step, torch.ones_like(step))
    nearest = torch.clamp(torch.round(pos), 1, n_bins - 1)
    return torch.abs(pos - nearest) * step


def dmel_tokenize(wav: torch.Tensor, cfg: MelConfig, lo: torch.Tensor, hi: torch.Tensor, n_bins: int,
                  bank: Optional[torch.Tensor] = None) -> torch.Tensor:
    """waveform -> uint8 codes, the whole path on the CPU."""
    return dmel_encode(log_mel(wav, cfg, bank), lo, hi, n_bins)


# --------------------------------------------------------------------------- #
# input side of the path, as the reference's data module does it on the host
# --------------------------------------------------------------------------- #
def peak_normalize(audio: torch.Tensor, target: float = 0.95) -> torch.Tensor:
    """``librosa.util.normalize(audio) * 0.95`` (reference dataset/lhotse_tts_dataset.py:32) for one 1-D utterance:
    x / max|x| in float32, then * 0.95; librosa treats a peak below the smallest normal float as 1."""
    x = audio.to(torch.float32)
    peak = x.abs().max()
    if peak.item() < torch.finfo(torch.float32).tiny:
        peak = torch.ones_like(peak)
    return (x / peak) * torch.tensor(target, dtype=torch.float32)


def log_mel_each(utterances, cfg: MelConfig, bank: Optional[torch.Tensor] = None, peak: Optional[float] = None):
    """The reference transform run on every utterance ALONE (what a batch size of 1 gives: reflect padding at
    the utterance's own ends, ``len // hop`` frames); optionally after the data module's peak normalisation.
    -> list of (n_mels, T_u) tensors."""
    out = []
    for u in utterances:
        x = u.reshape(-1).to(torch.float32)
        if peak is not None:
            x = peak_normalize(x, peak)
        out.append(log_mel(x[None, None, :], cfg, bank)[0])
    return out
