"""dMel bin quantiser, dequantiser and calibration.

The reference ships no such module — its quantiser is the learned
``DownsampleFiniteScalarQuantize`` (reference models/modules/dowmsample_fsq.py)
— so the arithmetic here is this repo's own spec.  The *API shape* mirrors that
reference class so a codec that swaps quantisers keeps its call sites:
``encode(z) -> codes``, ``decode(codes) -> z`` and ``forward(z) ->
Result(z, codes, latents)`` (dowmsample_fsq.py:12-16, :86, :124, :135), and
``DMelTokenizer.encode(audios, audio_lengths) -> (codes, code_lengths)`` mirrors
``VQGAN.encode`` (reference models/codec_lit_modules.py:462-466).

Per mel channel c with calibrated [lo_c, hi_c] and K bins, float32:

    code  = clamp(floor((x - lo_c) * (K / (hi_c - lo_c))), 0, K - 1)   uint8
    x_hat = lo_c + (code + 0.5) * ((hi_c - lo_c) / K)

Calibration is the per-channel min / max of log-mel over all valid frames
(t < audio_length // hop, the caller's mask rule at codec_lit_modules.py:176);
min and max are exact and order independent, so sharded calibration followed
by an all-reduce is bit-identical to a single-GPU pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Tuple

import torch
from torch import Tensor, nn

from . import plan as _plan
from .spectrogram import LogMelSpectrogram


@dataclass
class DMelResult:
    z: Tensor        # dequantised log-mel (B, n_mels, T) float32
    codes: Tensor    # (B, n_mels, T) uint8
    latents: Tensor  # the input log-mel


class DMelQuantizer(nn.Module):
    def __init__(self, n_mels: int, n_bins: int = 16):
        super().__init__()
        if not 1 <= n_bins <= 256:
            raise ValueError("n_bins must be in [1, 256] (codes are uint8)")
        self.n_mels = int(n_mels)
        self.n_bins = int(n_bins)
        # registered so calibration survives checkpoints
        self.register_buffer("lo", torch.full((n_mels,), float("inf"), dtype=torch.float32))
        self.register_buffer("hi", torch.full((n_mels,), float("-inf"), dtype=torch.float32))
        # derived tensors (scale, table, "is calibrated") are cached until the stats change,
        # so the steady-state encode path issues no extra kernels and no host sync
        self._derived = {}

    def _invalidate(self) -> None:
        self._derived = {}

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() move the buffers
        self._invalidate()
        lo32, hi32 = self.lo, self.hi
        out = super()._apply(fn, *args, **kwargs)
        # .half() / .bfloat16() / .double() on a parent module would cast the statistics too; the kernels read them
        # as float32 and the bin edges are defined in float32, so only the DEVICE follows such a call: the float32
        # values are carried over unrounded
        if self.lo.dtype != torch.float32:
            self.lo = lo32.to(self.lo.device)
            self.hi = hi32.to(self.hi.device)
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        self._invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    # -- statistics -----------------------------------------------------------
    def _derive(self) -> None:
        """scale, step and the device-side "is calibrated" flag: one launch on CUDA, the same arithmetic in torch ops
        on the CPU (where only the oracle-backed tests run)."""
        if "scale" in self._derived:
            return
        if self.lo.is_cuda:
            scale, step, flag = _plan.quantizer_derive(self.lo, self.hi, self.n_bins)
        else:
            width = self.hi - self.lo
            # a true float32 division K / width (scalar / tensor would be reciprocal-then-multiply: other bits)
            scale = torch.where(width > 0, torch.full_like(width, float(self.n_bins)) / width, torch.zeros_like(width))
            step = (width / float(self.n_bins)).contiguous()
            flag = torch.all(self.lo <= self.hi).to(torch.int32)
        self._derived.update(scale=scale, step=step, ready_flag=flag)

    @property
    def calibrated(self) -> bool:
        if "ready" not in self._derived:
            self._derive()
            self._derived["ready"] = bool(self._derived["ready_flag"].item())
        return self._derived["ready"]

    def reset_stats(self) -> None:
        self._invalidate()
        self.lo.fill_(float("inf"))
        self.hi.fill_(float("-inf"))

    def set_stats(self, lo: Tensor, hi: Tensor) -> None:
        self._invalidate()
        self.lo.copy_(lo.to(self.lo))
        self.hi.copy_(hi.to(self.hi))

    @torch.no_grad()
    def update_stats(self, mel: Tensor, mel_lengths: Optional[Tensor] = None) -> None:
        """Fold the min / max of a (B, n_mels, T) log-mel batch into lo / hi;
        frames at or past ``mel_lengths[b]`` are ignored."""
        self._check_channels(mel)
        self._invalidate()
        _plan.tensor_minmax(mel, mel_lengths, self.lo, self.hi)

    @torch.no_grad()
    def sync_stats(self, group=None) -> None:
        """All-reduce lo (MIN) and hi (MAX) across ranks: the one collective of
        the whole path (one message of 2 * n_mels floats, NCCL over NVLink when on GPUs)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        self._invalidate()
        # one collective: MIN over [lo, -hi] (negation is exact, so max(hi) = -min(-hi) bit for bit)
        both = torch.cat([self.lo, -self.hi])
        dist.all_reduce(both, op=dist.ReduceOp.MIN, group=group)
        self.lo.copy_(both[: self.n_mels])
        torch.neg(both[self.n_mels:], out=self.hi)

    def scale(self) -> Tensor:
        """K / (hi - lo) per channel, 0 where the channel is degenerate."""
        self._derive()
        return self._derived["scale"]

    def step(self) -> Tensor:
        """(hi - lo) / K per channel: the bin width the centre table is built from."""
        self._derive()
        return self._derived["step"]

    def host_stats(self):
        """(lo, scale) as CPU float32 tensors, copied from the device once per calibration (the
        host-buffer encode hands them to the library by host pointer)."""
        if "host_stats" not in self._derived:
            self._derived["host_stats"] = (self.lo.detach().to("cpu", torch.float32).contiguous(),
                                           self.scale().detach().to("cpu", torch.float32).contiguous())
        return self._derived["host_stats"]

    def table(self) -> Tensor:
        """(n_mels, K) bin centres, separate multiply and add (no FMA)."""
        if "table" not in self._derived:
            step = (self.hi - self.lo) / float(self.n_bins)
            k = torch.arange(self.n_bins, dtype=torch.float32, device=step.device) + 0.5
            prod = k[None, :] * step[:, None]
            self._derived["table"] = (self.lo[:, None] + prod).contiguous()
        return self._derived["table"]

    # -- codec API (names follow reference dowmsample_fsq.py:86/:124/:135) -----
    @torch.no_grad()
    def encode(self, z: Tensor, mel_lengths: Optional[Tensor] = None, *, check_after: bool = False) -> Tensor:
        """``mel_lengths``: valid frames per batch row; frames at or past it get code 0 (what the fused encode
        writes there) without their log-mel being read.  ``check_after``: queue the launch first and look at the calibration afterwards (the check reads a flag
        back from the device; in a job that has just all-reduced the statistics that read would otherwise sit
        between the collective and the launch with the GPU idle).  Raises all the same when uncalibrated."""
        if not check_after:
            self._check_ready()
        self._check_channels(z)
        pending = None
        if check_after and "ready" not in self._derived and self.lo.is_cuda:
            # the flag travels to pinned host memory BEFORE the launch is queued and is awaited after it: the host
            # learns it while the quantiser runs, and nothing sits behind the kernel on the stream
            if getattr(self, "_flag_host", None) is None:
                self._flag_host = torch.empty((), dtype=torch.int32).pin_memory()
            self._derive()
            self._flag_host.copy_(self._derived["ready_flag"], non_blocking=True)
            pending = torch.cuda.Event()
            pending.record(torch.cuda.current_stream(self.lo.device))
        codes = _plan.quantize(z, self.lo, self.scale(), self.n_bins, n_valid=mel_lengths)
        if pending is not None:
            pending.synchronize()
            self._derived["ready"] = bool(self._flag_host.item())
        if check_after:
            self._check_ready()
        return codes

    @torch.no_grad()
    def decode(self, indices: Tensor) -> Tensor:
        self._check_ready()
        return _plan.dequantize(indices, self.table())

    @torch.no_grad()
    def forward(self, z: Tensor) -> DMelResult:
        codes = self.encode(z)
        return DMelResult(z=self.decode(codes), codes=codes, latents=z)

    # -- helpers ----------------------------------------------------------------
    def _check_ready(self) -> None:
        if not self.calibrated:
            raise RuntimeError("DMelQuantizer has no calibration: call update_stats()/calibrate() or set_stats() first")

    def _check_channels(self, mel: Tensor) -> None:
        if mel.ndim != 3 or mel.shape[1] != self.n_mels:
            raise ValueError(f"expected (B, {self.n_mels}, T), got {tuple(mel.shape)}")


class DMelTokenizer(nn.Module):
    """waveform <-> dMel codes.  ``encode`` is the fused hot path: one kernel
    reads float32 samples and writes uint8 codes; log-mel never reaches HBM."""

    def __init__(self, sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=128,
                 n_bins=16, center=False, f_min=0.0, f_max=None):
        super().__init__()
        self.mel_transform = LogMelSpectrogram(sample_rate=sample_rate, n_fft=n_fft, win_length=win_length,
                                               hop_length=hop_length, n_mels=n_mels, center=center,
                                               f_min=f_min, f_max=f_max)
        self.quantizer = DMelQuantizer(n_mels, n_bins)
        self.hop_length = hop_length
        self.sample_rate = sample_rate

    def _plan(self, device: torch.device) -> _plan.Plan:
        return self.mel_transform.spectrogram.plan_for(device)

    @staticmethod
    def _flat_lengths(audio_lengths: Optional[Tensor]) -> Optional[Tensor]:
        # the reference data module yields audio_lengths as (1, B) int32
        # (dataset/lhotse_tts_dataset.py:60-65); sequence_mask squeezes it
        return None if audio_lengths is None else audio_lengths.reshape(-1)

    # -- calibration ----------------------------------------------------------
    @torch.no_grad()
    def update_stats(self, audios: Tensor, audio_lengths: Optional[Tensor] = None) -> None:
        """One calibration step on a batch of waveforms (fused: no log-mel tensor)."""
        q = self.quantizer
        q._invalidate()
        self._plan(audios.device).update_minmax(audios, self._flat_lengths(audio_lengths), q.lo, q.hi)

    @torch.no_grad()
    def update_stats_keep_mel(self, audios: Tensor, audio_lengths: Optional[Tensor] = None,
                              out: Optional[Tensor] = None) -> Tensor:
        """One calibration step that also returns the batch's log-mel (one launch), so that the encode pass of a
        calibrate-then-encode job can quantise the stored tensor instead of running the STFT again.  ``out``: where
        to write it (a contiguous (B, n_mels, T) float32 slice of a larger store)."""
        q = self.quantizer
        q._invalidate()
        return self._plan(audios.device).logmel_minmax(audios, self._flat_lengths(audio_lengths), q.lo, q.hi, out=out)

    def n_frames(self, n_samples: int) -> int:
        """Frames of an utterance of ``n_samples`` samples (reference utils/spectrogram.py:58-66: ``n // hop``)."""
        return int(n_samples) // self.hop_length

    @torch.no_grad()
    def calibrate(self, batches: Iterable, group=None) -> Tuple[Tensor, Tensor]:
        """Reset, scan ``batches`` (audios or (audios, audio_lengths)), then
        all-reduce across ranks.  Returns (lo, hi)."""
        self.quantizer.reset_stats()
        for item in batches:
            audios, lengths = (item if isinstance(item, (tuple, list)) else (item, None))
            self.update_stats(audios, lengths)
        self.quantizer.sync_stats(group)
        return self.quantizer.lo, self.quantizer.hi

    # -- codec API (reference VQGAN.encode / decode, codec_lit_modules.py:462-484)
    @torch.no_grad()
    def encode(self, audios: Tensor, audio_lengths: Optional[Tensor] = None, *, return_mel: bool = False,
               near_edge: Optional[Tensor] = None, edge_eps: float = 0.0):
        """-> (codes (B, n_mels, T) uint8, code_lengths or None[, log-mel])."""
        q = self.quantizer
        q._check_ready()
        lengths = self._flat_lengths(audio_lengths)
        out = self._plan(audios.device).encode(audios, lengths, q.lo, q.scale(), q.n_bins,
                                               return_logmel=return_mel, near_edge=near_edge, edge_eps=edge_eps)
        code_lengths = None if lengths is None else torch.div(lengths, self.hop_length, rounding_mode="floor")
        if return_mel:
            return out[0], code_lengths, out[1]
        return out, code_lengths

    @torch.no_grad()
    def encode_utterances(self, audios, audio_lengths: Optional[Tensor] = None, *, peak_normalize: Optional[float] = None,
                          return_mel: bool = False):
        """Encode every utterance as if the reference ran on it ALONE, without the host-side steps of the reference's
        data module (dataset/lhotse_tts_dataset.py:29-33 peak normalisation, :46-65 right-pad collate):

        * ``audios`` is a list of 1-D waveforms of different lengths (a ragged batch: they are packed back to back,
          no padding is stored or transformed), or a padded (B, L) / (B, 1, L) tensor with ``audio_lengths``;
        * the reflect padding happens at each utterance's own end and it has ``length // hop`` frames;
        * ``peak_normalize=0.95`` applies ``x / max|x| * 0.95`` per utterance on the GPU (one extra pass over the
          waveform, the normalised audio never exists in HBM).

        -> (codes (B, n_mels, T_max) uint8 zero past each utterance, code_lengths (B,)[, log-mel])."""
        q = self.quantizer
        q._check_ready()
        dev = q.lo.device
        plan = self._plan(dev)
        if isinstance(audios, (list, tuple)):
            lens = [int(a.numel()) for a in audios]
            # 16-byte aligned starts keep the bulk-copy path: round every utterance up to 4 samples inside the buffer
            starts, total = [], 0
            for n in lens:
                starts.append(total)
                total += (n + 3) // 4 * 4
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            for a, s0, n in zip(audios, starts, lens):
                flat[s0:s0 + n] = a.reshape(-1).to(device=dev, dtype=torch.float32)
            # row b = the first lengths[b] samples at offsets[b]; the slack up to offsets[b + 1] is never read
            offsets = torch.tensor(starts + [total], dtype=torch.int64, device=dev)
            lengths = torch.tensor(lens, dtype=torch.int32, device=dev)
            kw = dict(offsets=offsets, n_rows=len(lens), max_samples=max(lens), min_samples=min(lens), lengths=lengths)
            wav = flat
        else:
            if audio_lengths is None:
                raise ValueError("a padded batch needs audio_lengths to be treated utterance by utterance")
            lengths = self._flat_lengths(audio_lengths).to(dev)
            lens = lengths.tolist()
            kw = dict(lengths=lengths, own_length=True, min_samples=min(lens))
            wav = audios
        gain = None
        if peak_normalize is not None:
            gain = plan.peak_gain(wav, lengths=kw["lengths"], offsets=kw.get("offsets"), n_rows=kw.get("n_rows"),
                                  max_samples=kw.get("max_samples"), target=float(peak_normalize))
        out = plan.run(wav, row_gain=gain, lo=q.lo, scale=q.scale(), n_bins=q.n_bins, want_codes=True,
                       want_logmel=return_mel, mask_invalid=True, **kw)
        code_lengths = torch.div(kw["lengths"], self.hop_length, rounding_mode="floor")
        return (out["codes"], code_lengths, out["logmel"]) if return_mel else (out["codes"], code_lengths)

    @torch.no_grad()
    def encode_decode(self, audios: Tensor, audio_lengths: Optional[Tensor] = None) -> DMelResult:
        """The quantiser's forward fused with the transform, one launch: ``codes`` and ``z``, the bin
        centre of every code (== ``decode(codes, code_lengths)``, bit for bit).  ``latents`` (the
        pre-quantisation log-mel) never reaches HBM here and is None; ask ``encode(return_mel=True)`` for it."""
        q = self.quantizer
        q._check_ready()
        lengths = self._flat_lengths(audio_lengths)
        codes, mel_hat = self._plan(audios.device).encode_decode(audios, lengths, q.lo, q.scale(), q.step(), q.n_bins)
        return DMelResult(z=mel_hat, codes=codes, latents=None)

    @torch.no_grad()
    def encode_pcm16(self, audios: Tensor, audio_lengths: Optional[Tensor] = None):
        """``encode`` for int16 PCM on the device (value = sample / 32768): same codes as
        ``encode(audios.float() / 32768)``, half the waveform bytes.  -> (codes, code_lengths or None)."""
        q = self.quantizer
        q._check_ready()
        lengths = self._flat_lengths(audio_lengths)
        codes = self._plan(audios.device).encode_pcm16(audios, lengths, q.lo, q.scale(), q.n_bins)
        code_lengths = None if lengths is None else torch.div(lengths, self.hop_length, rounding_mode="floor")
        return codes, code_lengths

    @torch.no_grad()
    def encode_host(self, audios: Tensor, audio_lengths: Optional[Tensor] = None,
                    out: Optional[Tensor] = None) -> Tensor:
        """CPU (ideally pinned) waveforms in, CPU codes out, transfers pipelined natively.
        float32 waveforms, or int16 PCM (value = sample / 32768: half the PCIe bytes, same codes)."""
        q = self.quantizer
        q._check_ready()
        lo_h, scale_h = q.host_stats()
        return self._plan(q.lo.device).encode_host(audios, self._flat_lengths(audio_lengths), lo_h, scale_h,
                                                   q.n_bins, out)

    @torch.no_grad()
    def decode(self, indices: Tensor, feature_lengths: Optional[Tensor] = None) -> Tensor:
        """codes -> log-mel bin centres; frames past ``feature_lengths`` are zeroed
        like the reference's masked decode (codec_lit_modules.py:468-476)."""
        mel = self.quantizer.decode(indices)
        if feature_lengths is not None:
            t = torch.arange(mel.shape[2], device=mel.device)
            mel = mel * (t[None, None, :] < feature_lengths.reshape(-1, 1, 1).to(mel.device))
        return mel

    def forward(self, audios: Tensor, audio_lengths: Optional[Tensor] = None):
        return self.encode(audios, audio_lengths)
