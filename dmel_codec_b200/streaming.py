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
    """A lock-step batch of ``n_streams`` audio streams (rows advance together).

    Thin wrapper over the native ``dmel_stream_*`` entry points: the history buffer and the
    frame counters live in the library, and a push is one chunk copy plus one kernel launch.
    """

    def __init__(self, tokenizer: DMelTokenizer, n_streams: int = 1, capacity_samples: int = 1 << 16,
                 device: Optional[torch.device | str] = None):
        tokenizer.quantizer._check_ready()
        self.tok = tokenizer
        self.device = torch.device(device) if device is not None else tokenizer.quantizer.lo.device
        if self.device.type != "cuda":
            raise RuntimeError("DMelStreamEncoder needs the tokenizer on a CUDA device (no CPU fallback)")
        self.plan = tokenizer._plan(self.device)
        if tokenizer.mel_transform.center:
            raise NotImplementedError("streaming with center=True is not supported")
        self.n_streams = int(n_streams)
        self._lib = _native.load()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self._lib.dmel_stream_create(self.plan._handle, self.n_streams, int(capacity_samples),
                                                       ctypes.byref(handle)))
        self._handle = handle
        self._count = ctypes.c_longlong(0)
        self._count_ref = ctypes.byref(self._count)
        mt = tokenizer.mel_transform
        self._n_fft, self._hop = int(mt.n_fft), int(mt.hop_length)
        self._pad = (self._n_fft - self._hop) // 2
        self._seen = self._t_next = 0  # mirrors of the library's counters, for sizing the output

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                self._lib.dmel_stream_destroy(handle)
            except Exception:  # interpreter shutdown
                pass

    def reset(self) -> None:
        _native.check(self._lib.dmel_stream_reset(self._handle))
        self._seen = self._t_next = 0

    def frames_after(self, n: int) -> int:
        """Frames a push of ``n`` more samples per stream will emit (same arithmetic as
        ``dmel_stream_pending``, kept on the Python side so a push is a single call into the library)."""
        seen = self._seen + int(n)
        if seen + self._pad < self._n_fft or seen <= self._pad:
            return 0
        return max((seen + self._pad - self._n_fft) // self._hop + 1 - self._t_next, 0)

    # ------------------------------------------------------------------
    def push(self, chunk: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Append ``chunk`` ((n_streams, n) or (n,) float32, CUDA or CPU) and return
        the codes of every frame that became complete: (n_streams, n_mels, k) uint8, k >= 0.
        ``out``: an optional (n_streams, n_mels, frames_after(n)) uint8 CUDA tensor to write into
        (a latency-sensitive caller reuses its buffers)."""
        if chunk.ndim == 1:
            chunk = chunk[None, :]
        if chunk.shape[0] != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {chunk.shape[0]}")
        if chunk.dtype != torch.float32:
            chunk = chunk.float()
        if chunk.stride(1) != 1:
            chunk = chunk.contiguous()
        if chunk.is_cuda and chunk.device != self.device:
            chunk = chunk.to(self.device)
        n = chunk.shape[1]
        q = self.tok.quantizer
        count = self.frames_after(n)
        if out is None:
            out = torch.empty((self.n_streams, q.n_mels, count), dtype=torch.uint8, device=self.device)
        elif out.shape != (self.n_streams, q.n_mels, count) or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous uint8 tensor of shape {(self.n_streams, q.n_mels, count)}")
        _native.check(self._lib.dmel_stream_push(
            self._handle, chunk.data_ptr(), n, chunk.stride(0) if self.n_streams > 1 else n, q.lo.data_ptr(),
            q.scale().data_ptr(), q.n_bins, out.data_ptr(), count, self._count_ref,
            torch.cuda.current_stream(self.device).cuda_stream))
        self._seen += n
        self._t_next += count
        return out

    # -- zero-copy chunks: the producer writes straight into the history buffer -----------------------------
    def input_view(self, n: int) -> torch.Tensor:
        """(n_streams, n) float32 CUDA view of where the next ``n`` samples of every stream belong (a window of the
        library's history buffer).  Fill it (a kernel of the producer, or a host-to-device copy straight into it),
        then ``commit(n)``: one kernel launch per chunk, no copy of the chunk."""
        if not getattr(self, "_bound", False):
            q = self.tok.quantizer
            self._lo, self._scale = q.lo, q.scale()  # kept alive: the library holds their addresses
            _native.check(self._lib.dmel_stream_bind(self._handle, self._lo.data_ptr(), self._scale.data_ptr(), q.n_bins,
                                                     torch.cuda.current_stream(self.device).cuda_stream))
            self._where, self._stride = ctypes.c_void_p(), ctypes.c_longlong()
            self._bound, self._views = True, {}
        _native.check(self._lib.dmel_stream_input(self._handle, n, ctypes.byref(self._where), ctypes.byref(self._stride)))
        key = (self._where.value, n)
        view = self._views.get(key)
        if view is None:  # wrap the device window as a tensor once per distinct (address, length)
            if len(self._views) > 4096:
                self._views.clear()
            view = _device_view(self._where.value, self.n_streams, n, self._stride.value, self.device)
            self._views[key] = view
        return view

    def commit(self, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Encode every frame completed by the ``n`` samples written into ``input_view(n)``."""
        count = self.frames_after(n)
        if out is None:
            out = torch.empty((self.n_streams, self.tok.quantizer.n_mels, count), dtype=torch.uint8, device=self.device)
        elif out.shape[2] != count or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous uint8 tensor with {count} frames")
        _native.check(self._lib.dmel_stream_commit(self._handle, n, out.data_ptr(), count, self._count_ref))
        self._seen += n
        self._t_next += count
        return out

    def flush(self) -> torch.Tensor:
        """End of stream: emit the remaining frames (they use the reference's right-edge
        reflection) and reset.  Total frames over the stream's life = n_samples // hop."""
        q = self.tok.quantizer
        count = max(self._lib.dmel_stream_pending(self._handle, 0, 1), 0)
        codes = torch.empty((self.n_streams, q.n_mels, count), dtype=torch.uint8, device=self.device)
        _native.check(self._lib.dmel_stream_flush(
            self._handle, q.lo.data_ptr(), q.scale().data_ptr(), q.n_bins, codes.data_ptr(), count,
            self._count_ref, torch.cuda.current_stream(self.device).cuda_stream))
        self._seen = self._t_next = 0  # the library resets the stream after a flush
        return codes


class _DevicePtr:
    """Minimal ``__cuda_array_interface__`` carrier: lets torch wrap a window of library-owned device memory."""

    def __init__(self, ptr: int, shape, strides_bytes):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": tuple(strides_bytes)}


def _device_view(ptr: int, rows: int, n: int, row_stride: int, device: torch.device) -> torch.Tensor:
    with torch.cuda.device(device):
        return torch.as_tensor(_DevicePtr(ptr, (rows, n), (row_stride * 4, 4)), device=device)
