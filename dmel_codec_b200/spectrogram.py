"""Drop-in replacements for ``dmel_codec.utils.spectrogram.LinearSpectrogram``
and ``LogMelSpectrogram`` (reference dmel_codec/utils/spectrogram.py:7-127).

Same constructor keywords, defaults, attributes and call signature; the forward
pass is one fused sm_100a kernel (``csrc/logmel_kernel.cuh``) reached through
the C ABI instead of the reference's pad / torch.stft / sqrt / matmul / log op
chain.  Output: ``(B, n_mels, T)`` float32 on the input's CUDA device, no
autograd graph (the reference's call sites run under ``torch.no_grad``,
models/codec_lit_modules.py:170).  CPU tensors raise: there is no fallback.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor, nn

from .plan import Plan


class LinearSpectrogram(nn.Module):
    """Despite its name the reference class returns the log-mel spectrogram
    (utils/spectrogram.py:78-81); so does this one."""

    def __init__(self, n_fft=2048, win_length=2048, hop_length=512, center=False, num_mels=128,
                 f_min=0, f_max=None, sample_rate=44100, mode="reflect"):
        super().__init__()
        self.n_fft = n_fft
        self.win_length = win_length
        self.hop_length = hop_length
        self.center = center
        self.mode = mode
        self.f_min = f_min
        self.f_max = f_max
        self.num_mels = num_mels
        self.sample_rate = sample_rate
        # per-device native plans; like the reference's mel_basis_cache /
        # hann_window_cache (:32-33) these are plain dicts, not buffers, so they
        # carry no state_dict entries and ignore .to()/.half() on the parent
        self._plans: Dict[str, Plan] = {}

    def plan_for(self, device: torch.device) -> Plan:
        if self.mode != "reflect":
            raise NotImplementedError(f"pad mode {self.mode!r}: only 'reflect' (the reference's only caller value) is built")
        key = f"{self.n_fft}_{self.num_mels}_{self.sample_rate}_{self.hop_length}_{self.win_length}_{self.f_min}_{self.f_max}_{self.center}_{device}"
        plan = self._plans.get(key)
        if plan is None:
            plan = Plan(sample_rate=self.sample_rate, n_fft=self.n_fft, win_length=self.win_length,
                        hop_length=self.hop_length, n_mels=self.num_mels, f_min=self.f_min, f_max=self.f_max,
                        center=self.center, device=device)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def forward(self, y: Tensor) -> Tensor:
        if not y.is_cuda:
            raise RuntimeError("dmel_codec_b200.LinearSpectrogram needs a CUDA tensor (no CPU fallback); "
                               f"got device {y.device}")
        return self.plan_for(y.device).logmel(y)


class LogMelSpectrogram(nn.Module):
    def __init__(self, sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=128,
                 center=False, f_min=0.0, f_max=None):
        super().__init__()
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.win_length = win_length
        self.hop_length = hop_length
        self.center = center
        self.n_mels = n_mels
        self.f_min = f_min
        self.f_max = f_max or float(sample_rate // 2)  # reference :105
        # the reference forwards the RAW f_max, not self.f_max (:114)
        self.spectrogram = LinearSpectrogram(n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                                             center=center, num_mels=n_mels, f_min=f_min, f_max=f_max,
                                             sample_rate=sample_rate, mode="reflect")

    @torch.no_grad()
    def forward(self, x: Tensor, return_linear: bool = False, sample_rate: Optional[int] = None) -> Tensor:
        # `return_linear` is accepted and ignored, as in the reference (:119-127)
        if sample_rate is not None and sample_rate != self.sample_rate:
            import torchaudio.functional as AF  # off the hot path; no shipped caller passes sample_rate
            x = AF.resample(x, orig_freq=sample_rate, new_freq=self.sample_rate)
        return self.spectrogram(x)

    @torch.no_grad()
    def masked(self, audios: Tensor, audio_lengths: Tensor, dtype: torch.dtype = torch.float32):
        """``(mels, mel_lengths)`` as ``VQGAN.encode_unquantized`` builds them before the encoder
        (reference models/codec_lit_modules.py:486-507): log-mel cast to ``dtype`` with the frames at
        or past ``audio_lengths // hop_length`` zeroed — transform, cast and ``sequence_mask``
        multiply in one kernel launch, padded frames never computed.  The caller's group view
        ``mels.view(B * G, n_mels // G, T)`` needs no copy."""
        if not audios.is_cuda:
            raise RuntimeError("dmel_codec_b200.LogMelSpectrogram needs a CUDA tensor (no CPU fallback); "
                               f"got device {audios.device}")
        mels = self.spectrogram.plan_for(audios.device).logmel_masked(audios, audio_lengths, dtype)
        return mels, audio_lengths // self.hop_length

    @torch.no_grad()
    def with_quality(self, audios: Tensor):
        """``(mels, quality)`` of ``VQGAN.training_step`` (reference models/codec_lit_modules.py:171-174):
        ``quality = ((mels.mean(-1) > -8).sum(-1) - 90) / 10`` with the per-channel time sums
        accumulated by the same launch that writes the mel (one transform call serves both the
        reference's identically configured ``encode_mel_transform`` and ``gt_mel_transform``)."""
        if not audios.is_cuda:
            raise RuntimeError("dmel_codec_b200.LogMelSpectrogram needs a CUDA tensor (no CPU fallback); "
                               f"got device {audios.device}")
        plan = self.spectrogram.plan_for(audios.device)
        b = audios.shape[0]
        sums = torch.zeros((b, self.n_mels), dtype=torch.float32, device=audios.device)
        mels = plan.logmel_masked(audios, None, torch.float32, row_sum=sums)
        quality = ((sums / mels.shape[-1] > -8).sum(-1) - 90) / 10
        return mels, quality.unsqueeze(-1)

