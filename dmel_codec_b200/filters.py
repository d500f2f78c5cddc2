"""Host-side constants of the log-mel transform: Slaney mel filterbank, STFT
window and the banded (sparse) form of the filterbank the CUDA kernels read.

The reference builds these lazily inside ``LinearSpectrogram.forward``
(reference utils/spectrogram.py:43-56) from ``librosa.filters.mel`` (librosa
0.10.2.post1, an un-vendored dependency) and ``torch.hann_window``.  Here they
are built once per plan, in float64 on the host, and rounded to float32 at the
same points librosa rounds.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

_LIN_HZ_PER_MEL = 200.0 / 3.0
_KNEE_HZ = 1000.0
_KNEE_MEL = _KNEE_HZ / _LIN_HZ_PER_MEL
_LOG_STEP = np.log(6.4) / 27.0


def hz_to_mel(hz):
    """Slaney mel scale (linear to 1 kHz, log above) — librosa's ``htk=False``."""
    hz = np.atleast_1d(np.asarray(hz, dtype=np.float64))
    mel = hz / _LIN_HZ_PER_MEL
    hi = hz >= _KNEE_HZ
    mel[hi] = _KNEE_MEL + np.log(hz[hi] / _KNEE_HZ) / _LOG_STEP
    return mel


def mel_to_hz(mel):
    mel = np.atleast_1d(np.asarray(mel, dtype=np.float64))
    hz = mel * _LIN_HZ_PER_MEL
    hi = mel >= _KNEE_MEL
    hz[hi] = _KNEE_HZ * np.exp(_LOG_STEP * (mel[hi] - _KNEE_MEL))
    return hz


def mel_filterbank(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0,
                   f_max: Optional[float] = None) -> np.ndarray:
    """(n_mels, n_fft//2+1) float32, equal to
    ``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` with its defaults
    (call site: reference utils/spectrogram.py:45-51).  ``f_max=None`` means
    Nyquist, which is what reaches librosa because the reference forwards the
    raw argument (utils/spectrogram.py:114)."""
    if f_max is None:
        f_max = sample_rate / 2.0
    n_freq = n_fft // 2 + 1
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sample_rate)  # librosa.fft_frequencies; (n_freq,)
    assert freqs.shape[0] == n_freq
    mel_lo, mel_hi = hz_to_mel(f_min)[0], hz_to_mel(f_max)[0]
    corners = mel_to_hz(np.linspace(mel_lo, mel_hi, n_mels + 2))
    span = corners[1:] - corners[:-1]  # (M+1,)
    offset = corners[:, None] - freqs[None, :]  # (M+2, F)
    up = -offset[:-2] / span[:-1, None]
    down = offset[2:] / span[1:, None]
    tri = np.maximum(0.0, np.minimum(up, down)).astype(np.float32)  # librosa stores fp32 here
    tri *= (2.0 / (corners[2:] - corners[:-2]))[:, None]  # fp64 factor into the fp32 array
    return tri


def stft_window(win_length: int, n_fft: int) -> np.ndarray:
    """Periodic Hann (``torch.hann_window``, reference utils/spectrogram.py:53)
    centre-padded to ``n_fft`` as ``torch.stft`` does for a short window."""
    w = torch.hann_window(win_length, dtype=torch.float32).numpy()
    if win_length > n_fft:
        raise ValueError("win_length must not exceed n_fft")
    out = np.zeros(n_fft, dtype=np.float32)
    left = (n_fft - win_length) // 2
    out[left:left + win_length] = w
    return out


@dataclass
class BandedBank:
    """Filterbank rows cut down to their non-zero span.

    ``start[m]`` first frequency bin of channel m, ``count[m]`` bins in the span
    (zeros inside a span are kept so spans stay contiguous), ``offset[m]`` where
    the span's weights begin in ``weights``.  Each span is padded with zero
    weights to a multiple of 4 and starts 16-byte aligned so the kernel can
    fetch four weights per shared-memory load.
    """
    start: np.ndarray  # int32 (M,)
    count: np.ndarray  # int32 (M,)  padded to a multiple of 4
    offset: np.ndarray  # int32 (M,)
    weights: np.ndarray  # float32 (nnz_padded,)
    n_freq: int

    @property
    def nnz(self) -> int:
        return int(self.weights.size)


def band_filterbank(bank: np.ndarray) -> BandedBank:
    n_mels, n_freq = bank.shape
    start = np.zeros(n_mels, np.int32)
    count = np.zeros(n_mels, np.int32)
    offset = np.zeros(n_mels, np.int32)
    chunks = []
    cursor = 0
    for m in range(n_mels):
        nz = np.flatnonzero(bank[m])
        if nz.size == 0:
            s, c = 0, 0
        else:
            s, c = int(nz[0]), int(nz[-1] - nz[0] + 1)
        c4 = (c + 3) // 4 * 4
        row = np.zeros(c4, np.float32)
        row[:c] = bank[m, s:s + c]
        start[m], count[m], offset[m] = s, c4, cursor
        chunks.append(row)
        cursor += c4
    weights = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
    return BandedBank(start, count, offset, weights, n_freq)
