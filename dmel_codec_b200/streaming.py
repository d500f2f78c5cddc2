"""Chunked (streaming) dMel encode — BASELINE configs[3]: batch 1, 80 ms chunks.

The reference has no streaming mode; the contract here is that the codes
emitted chunk by chunk, concatenated, are bit-identical to ``DMelTokenizer.encode``
on the whole waveform (same kernel, same per-frame arithmetic).

Frame t needs samples [t*hop - pad, t*hop - pad + n_fft) with
pad = (n_fft - hop)//2, so a frame is emitted as soon as its last tap has
arrived (2.5 hops of algorithmic look-ahead at n_fft 1024 / hop 256) and the
encoder keeps ``n_fft - hop + pad``-ish samples of history per stream.  The
frames that need the right-edge reflection are emitted by ``flush()``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _native
from .quantizer import DMelTokenizer


class DMelStreamEncoder:
    """A lock-step batch of ``n_streams`` audio streams (rows advance together)."""

    def __init__(self, tokenizer: DMelTokenizer, n_streams: int = 1, capacity_samples: int = 1 << 16,
                 device: Optional[torch.device | str] = None):
        tokenizer.quantizer._check_ready()
        self.tok = tokenizer
        self.device = torch.device(device) if device is not None else tokenizer.quantizer.lo.device
        if self.device.type != "cuda":
            raise RuntimeError("DMelStreamEncoder needs the tokenizer on a CUDA device (no CPU fallback)")
        self.plan = tokenizer._plan(self.device)
        mt = tokenizer.mel_transform
        if mt.center:
            raise NotImplementedError("streaming with center=True is not supported")
        self.n_fft, self.hop = mt.n_fft, mt.hop_length
        self.pad = (self.n_fft - self.hop) // 2
        self.n_streams = int(n_streams)
        self.capacity = max(int(capacity_samples), 4 * self.n_fft) // 4 * 4
        self.buf = torch.zeros((self.n_streams, self.capacity), dtype=torch.float32, device=self.device)
        self.reset()

    def reset(self) -> None:
        self.base = 0     # virtual sample index of buf[:, 0]
        self.seen = 0     # samples received per stream
        self.t_next = 0   # next frame to emit

    # ------------------------------------------------------------------
    def push(self, chunk: torch.Tensor) -> torch.Tensor:
        """Append ``chunk`` ((n_streams, n) or (n,) float32, CUDA or CPU) and return
        the codes of every frame that became complete: (n_streams, n_mels, k) uint8, k >= 0."""
        if chunk.ndim == 1:
            chunk = chunk[None, :]
        if chunk.shape[0] != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {chunk.shape[0]}")
        n = chunk.shape[1]
        if self.seen - self.base + n > self.capacity:
            self._compact(n)
        at = self.seen - self.base
        self.buf[:, at:at + n].copy_(chunk, non_blocking=True)
        self.seen += n
        t_end = 0
        if self.seen + self.pad >= self.n_fft and self.seen > self.pad:
            t_end = (self.seen + self.pad - self.n_fft) // self.hop + 1
        return self._emit(t_end, self.seen)

    def flush(self) -> torch.Tensor:
        """End of stream: emit the remaining frames (they use the reference's right-edge
        reflection) and reset.  Total frames over the stream's life = n_samples // hop."""
        if self.seen <= self.pad:
            raise ValueError(f"stream of {self.seen} samples is shorter than the reflect pad {self.pad}")
        out = self._emit(self.plan.num_frames(self.seen), self.seen)
        self.reset()
        return out

    # ------------------------------------------------------------------
    def _emit(self, t_end: int, n_samples: int) -> torch.Tensor:
        count = t_end - self.t_next
        q = self.tok.quantizer
        codes = torch.empty((self.n_streams, q.n_mels, max(count, 0)), dtype=torch.uint8, device=self.device)
        if count > 0:
            _native.check(_native.load().dmel_encode_frames_u8(
                self.plan._handle, self.buf.data_ptr(), self.n_streams, self.capacity, self.base, n_samples,
                self.t_next, count, q.lo.data_ptr(), q.scale().data_ptr(), q.n_bins, codes.data_ptr(), None,
                torch.cuda.current_stream(self.device).cuda_stream))
            self.t_next = t_end
        return codes

    def _compact(self, incoming: int) -> None:
        """Drop samples no future frame needs; keep the new base 16-byte aligned."""
        keep_from = max(0, self.t_next * self.hop - self.pad) // 4 * 4
        keep_from = max(keep_from, self.base)
        live = self.seen - keep_from
        if live + incoming > self.capacity:
            raise ValueError(f"chunk of {incoming} samples does not fit a stream buffer of {self.capacity}")
        if keep_from > self.base:
            src = self.buf[:, keep_from - self.base:self.seen - self.base]
            self.buf[:, :live].copy_(src.clone() if live > keep_from - self.base else src)
            self.base = keep_from
