"""Python handle on a native ``dmel_plan`` plus the tensor-level call wrappers.

A plan fixes the transform geometry (n_fft, hop, mel filterbank, window) on one
GPU.  The wrappers only do what torch is here for: allocate outputs, hand over
raw device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _native, filters


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must be a CUDA tensor: dmel_codec_b200 runs on the GPU only (no CPU fallback); got device {t.device}")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_stat(t: torch.Tensor, what: str, device: torch.device, n: int) -> None:
    """Per-channel statistics and tables go to the kernels as raw float32 pointers: a tensor on another device,
    of another dtype or too short would be read as garbage (or fault the context), so check before the call."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != device:
        raise ValueError(f"{what} must be a CUDA tensor on {device} (the audio's device); got "
                         f"{getattr(t, 'device', type(t))}. Move the module with .to(device) first")
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() < n:
        raise ValueError(f"{what} must be a contiguous float32 tensor of at least {n} elements, "
                         f"got {t.dtype} {tuple(t.shape)} (contiguous={t.is_contiguous()})")


def as_rows(wav: torch.Tensor) -> torch.Tensor:
    """(B, L) or (B, 1, L) -> (B, L) float32 with unit stride along L, the two
    input layouts the reference accepts (utils/spectrogram.py:60-62)."""
    if wav.ndim == 3:
        if wav.shape[1] != 1:
            raise ValueError(f"expected mono audio (B, 1, L), got {tuple(wav.shape)}")
        wav = wav[:, 0, :]
    elif wav.ndim != 2:
        raise ValueError(f"expected audio of shape (B, L) or (B, 1, L), got {tuple(wav.shape)}")
    if wav.dtype != torch.float32:
        wav = wav.float()
    if wav.shape[0] > 1 and (wav.stride(1) != 1 or wav.stride(0) < wav.shape[1]):
        wav = wav.contiguous()
    elif wav.stride(1) != 1:
        wav = wav.contiguous()
    return wav


def as_pcm_rows(wav: torch.Tensor) -> torch.Tensor:
    """(B, L) or (B, 1, L) int16 -> (B, L) int16 with unit stride along L."""
    if wav.dtype != torch.int16:
        raise ValueError(f"expected int16 PCM, got {wav.dtype}")
    if wav.ndim == 3:
        if wav.shape[1] != 1:
            raise ValueError(f"expected mono audio (B, 1, L), got {tuple(wav.shape)}")
        wav = wav[:, 0, :]
    elif wav.ndim != 2:
        raise ValueError(f"expected audio of shape (B, L) or (B, 1, L), got {tuple(wav.shape)}")
    if wav.stride(1) != 1 or (wav.shape[0] > 1 and wav.stride(0) < wav.shape[1]):
        wav = wav.contiguous()
    return wav


class Plan:
    def __init__(self, *, sample_rate: int, n_fft: int, win_length: int, hop_length: int, n_mels: int,
                 f_min: float = 0.0, f_max: Optional[float] = None, center: bool = False,
                 device: torch.device | str | int = "cuda"):
        self._handle = ctypes.c_void_p()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"dmel_codec_b200 plans live on CUDA devices only, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_fft, self.hop_length, self.n_mels, self.center = int(n_fft), int(hop_length), int(n_mels), bool(center)
        self.mel_basis = np.ascontiguousarray(
            filters.mel_filterbank(sample_rate, n_fft, n_mels, f_min, f_max), dtype=np.float32)
        self.window = np.ascontiguousarray(filters.stft_window(win_length, n_fft), dtype=np.float32)
        lib = _native.load()
        with torch.cuda.device(self.device):
            _native.check(lib.dmel_plan_create(
                self.n_fft, self.hop_length, self.n_mels, int(self.center),
                self.mel_basis.ctypes.data_as(ctypes.c_void_p), self.window.ctypes.data_as(ctypes.c_void_p),
                ctypes.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            self._handle = None  # at interpreter shutdown module globals (ctypes, _native) may already be gone
            try:
                _native.load().dmel_plan_destroy(h)
            except Exception:
                pass

    def describe(self) -> dict:
        """Launch configuration the library chose for this geometry (diagnostics)."""
        import json
        buf = ctypes.create_string_buffer(512)
        _native.check(_native.load().dmel_plan_describe(self._handle, buf, len(buf)))
        return json.loads(buf.value.decode())

    def num_frames(self, n_samples: int) -> int:
        return int(_native.load().dmel_plan_num_frames(self._handle, int(n_samples)))

    # -- waveform -> log-mel ------------------------------------------------
    def logmel(self, wav: torch.Tensor) -> torch.Tensor:
        _require_cuda(wav, "audio")
        rows = as_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        out = torch.empty((b, self.n_mels, t), dtype=torch.float32, device=rows.device)
        _native.check(_native.load().dmel_logmel_f32(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, out.data_ptr(),
            _stream_ptr(rows.device)))
        return out

    def logmel_masked(self, wav: torch.Tensor, lengths: Optional[torch.Tensor],
                      dtype: torch.dtype = torch.float32, row_sum: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Log-mel in ``dtype`` (float32 or bfloat16) with frames at or past ``lengths // hop``
        zeroed, in one launch (``dmel_logmel_masked``)."""
        _require_cuda(wav, "audio")
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError(f"dtype must be torch.float32 or torch.bfloat16, got {dtype}")
        rows = as_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        if row_sum is not None:
            _require_stat(row_sum, "row_sum", rows.device, b * self.n_mels)
        out = torch.empty((b, self.n_mels, t), dtype=dtype, device=rows.device)
        len_ptr = self._lengths_ptr(lengths, b, rows.device)
        _native.check(_native.load().dmel_logmel_masked(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, len_ptr[0],
            1 if dtype == torch.bfloat16 else 0, out.data_ptr(),
            row_sum.data_ptr() if row_sum is not None else None, _stream_ptr(rows.device)))
        return out

    # -- calibration pass ---------------------------------------------------
    def update_minmax(self, wav: torch.Tensor, lengths: Optional[torch.Tensor], run_min: torch.Tensor,
                      run_max: torch.Tensor) -> None:
        _require_cuda(wav, "audio")
        rows = as_rows(wav)
        b, n = rows.shape
        self._frames_or_raise(n)
        _require_stat(run_min, "run_min", rows.device, self.n_mels)
        _require_stat(run_max, "run_max", rows.device, self.n_mels)
        len_ptr = self._lengths_ptr(lengths, b, rows.device)
        _native.check(_native.load().dmel_minmax_f32(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, len_ptr[0],
            run_min.data_ptr(), run_max.data_ptr(), _stream_ptr(rows.device)))

    def logmel_minmax(self, wav: torch.Tensor, lengths: Optional[torch.Tensor], run_min: torch.Tensor,
                      run_max: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Log-mel of every frame, and the running per-channel min / max over the valid ones, one launch.
        ``out``: a contiguous float32 (B, n_mels, T) tensor to write into (e.g. a slice of a shard-sized store)."""
        _require_cuda(wav, "audio")
        rows = as_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        _require_stat(run_min, "run_min", rows.device, self.n_mels)
        _require_stat(run_max, "run_max", rows.device, self.n_mels)
        if out is None:
            out = torch.empty((b, self.n_mels, t), dtype=torch.float32, device=rows.device)
        elif (tuple(out.shape) != (b, self.n_mels, t) or out.dtype != torch.float32 or out.device != rows.device
              or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous float32 {(b, self.n_mels, t)} tensor on {rows.device}, got "
                             f"{out.dtype} {tuple(out.shape)} on {out.device}")
        len_ptr = self._lengths_ptr(lengths, b, rows.device)
        _native.check(_native.load().dmel_logmel_minmax_f32(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, len_ptr[0], out.data_ptr(),
            run_min.data_ptr(), run_max.data_ptr(), _stream_ptr(rows.device)))
        return out

    # -- waveform -> codes, fused --------------------------------------------
    def encode(self, wav: torch.Tensor, lengths: Optional[torch.Tensor], lo: torch.Tensor, scale: torch.Tensor,
               n_bins: int, *, return_logmel: bool = False, near_edge: Optional[torch.Tensor] = None,
               edge_eps: float = 0.0):
        _require_cuda(wav, "audio")
        rows = as_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        _require_stat(lo, "lo", rows.device, self.n_mels)
        _require_stat(scale, "scale", rows.device, self.n_mels)
        codes = torch.empty((b, self.n_mels, t), dtype=torch.uint8, device=rows.device)
        logmel = torch.empty((b, self.n_mels, t), dtype=torch.float32, device=rows.device) if return_logmel else None
        len_ptr = self._lengths_ptr(lengths, b, rows.device)
        _native.check(_native.load().dmel_encode_u8(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, len_ptr[0],
            lo.data_ptr(), scale.data_ptr(), int(n_bins), codes.data_ptr(),
            logmel.data_ptr() if logmel is not None else None,
            near_edge.data_ptr() if near_edge is not None else None, float(edge_eps),
            _stream_ptr(rows.device)))
        return (codes, logmel) if return_logmel else codes

    def encode_decode(self, wav: torch.Tensor, lengths: Optional[torch.Tensor], lo: torch.Tensor, scale: torch.Tensor,
                      step: torch.Tensor, n_bins: int):
        """(codes, mel_hat): codes as ``encode`` and the bin centre of every code, one launch."""
        _require_cuda(wav, "audio")
        rows = as_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        for name, stat in (("lo", lo), ("scale", scale), ("step", step)):
            _require_stat(stat, name, rows.device, self.n_mels)
        codes = torch.empty((b, self.n_mels, t), dtype=torch.uint8, device=rows.device)
        mel_hat = torch.empty((b, self.n_mels, t), dtype=torch.float32, device=rows.device)
        len_ptr = self._lengths_ptr(lengths, b, rows.device)
        _native.check(_native.load().dmel_encode_decode_u8(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, len_ptr[0],
            lo.data_ptr(), scale.data_ptr(), step.data_ptr(), int(n_bins), codes.data_ptr(), mel_hat.data_ptr(),
            _stream_ptr(rows.device)))
        return codes, mel_hat

    def encode_pcm16(self, wav: torch.Tensor, lengths: Optional[torch.Tensor], lo: torch.Tensor,
                     scale: torch.Tensor, n_bins: int) -> torch.Tensor:
        """int16 PCM (value = sample / 32768) -> codes; bit-identical to ``encode(wav.float() / 32768)``."""
        _require_cuda(wav, "audio")
        rows = as_pcm_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        _require_stat(lo, "lo", rows.device, self.n_mels)
        _require_stat(scale, "scale", rows.device, self.n_mels)
        codes = torch.empty((b, self.n_mels, t), dtype=torch.uint8, device=rows.device)
        len_ptr = self._lengths_ptr(lengths, b, rows.device)
        _native.check(_native.load().dmel_encode_pcm16_u8(
            self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n, len_ptr[0],
            lo.data_ptr(), scale.data_ptr(), int(n_bins), codes.data_ptr(), _stream_ptr(rows.device)))
        return codes

    def encode_host(self, wav: torch.Tensor, lengths: Optional[torch.Tensor], lo: torch.Tensor,
                    scale: torch.Tensor, n_bins: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host tensors in, host tensor out; copies and kernels are pipelined
        inside the library (``dmel_encode_host_u8``)."""
        if wav.is_cuda:
            raise ValueError("encode_host takes CPU tensors; use encode() for CUDA tensors")
        pcm = wav.dtype == torch.int16  # int16 PCM: value = sample / 32768, half the bytes over PCIe
        rows = as_pcm_rows(wav) if pcm else as_rows(wav)
        b, n = rows.shape
        t = self._frames_or_raise(n)
        if out is None:
            out = torch.empty((b, self.n_mels, t), dtype=torch.uint8, pin_memory=True)
        elif (out.is_cuda or out.dtype != torch.uint8 or tuple(out.shape) != (b, self.n_mels, t)
              or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous CPU uint8 tensor of shape {(b, self.n_mels, t)}, got "
                             f"{out.dtype} {tuple(out.shape)} on {out.device}")
        lo_h = lo.detach().to("cpu", torch.float32).contiguous()
        sc_h = scale.detach().to("cpu", torch.float32).contiguous()
        len_h = None
        if lengths is not None:
            len_h = lengths.detach().to("cpu", torch.int32).reshape(-1).contiguous()
            if len_h.numel() != b:
                raise ValueError(f"lengths has {len_h.numel()} entries for a batch of {b}")
        with torch.cuda.device(self.device):
            fn = _native.load().dmel_encode_host_pcm16_u8 if pcm else _native.load().dmel_encode_host_u8
            _native.check(fn(
                self._handle, rows.data_ptr(), b, n, rows.stride(0) if b > 1 else n,
                len_h.data_ptr() if len_h is not None else None, lo_h.data_ptr(), sc_h.data_ptr(),
                int(n_bins), out.data_ptr()))
        return out

    # -- the general call: every input layout, every output ---------------------
    def run(self, wav: torch.Tensor, *, lengths: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
            n_rows: Optional[int] = None, max_samples: Optional[int] = None, min_samples: Optional[int] = None,
            own_length: bool = False, row_gain: Optional[torch.Tensor] = None,
            lo: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None, step: Optional[torch.Tensor] = None,
            n_bins: int = 0, want_codes: bool = False, want_mel_hat: bool = False, want_logmel: bool = False,
            logmel_dtype: torch.dtype = torch.float32, mask_invalid: bool = False,
            run_min: Optional[torch.Tensor] = None, run_max: Optional[torch.Tensor] = None) -> dict:
        """``dmel_run``: padded ``wav`` (B, L) / (B, 1, L), or a flat ragged buffer with ``offsets`` (B + 1 int64,
        CUDA; then ``n_rows``, ``max_samples`` and ``min_samples`` describe the rows).  Returns the requested
        outputs by name: ``codes``, ``mel_hat``, ``logmel``."""
        _require_cuda(wav, "audio")
        pcm = wav.dtype == torch.int16
        dev = wav.device
        if offsets is not None:
            flat = wav.reshape(-1)
            if flat.stride(0) != 1:
                flat = flat.contiguous()
            if not pcm and flat.dtype != torch.float32:
                flat = flat.float()
            if n_rows is None or max_samples is None or min_samples is None:
                raise ValueError("a ragged batch needs n_rows, max_samples and min_samples (host-side numbers)")
            off = offsets.to(device=dev, dtype=torch.int64).contiguous()
            if off.numel() != n_rows + 1:
                raise ValueError(f"offsets has {off.numel()} entries for {n_rows} rows")
            if off.data_ptr() != offsets.data_ptr():
                off.record_stream(torch.cuda.current_stream(dev))
            rows, b, n, stride = flat, int(n_rows), int(max_samples), int(max_samples)
        else:
            rows = as_pcm_rows(wav) if pcm else as_rows(wav)
            b, n = rows.shape
            stride = rows.stride(0) if b > 1 else n
            off = None
            if own_length and min_samples is None:
                raise ValueError("own_length needs min_samples: the shortest utterance, known on the host")
        t = self._frames_or_raise(n)
        io = _native.DmelIO()
        io.struct_size = ctypes.sizeof(_native.DmelIO)
        io.wav_dev, io.wav_is_pcm16 = rows.data_ptr(), int(pcm)
        io.n_rows, io.n_samples, io.row_stride = b, n, stride
        io.offsets_dev = off.data_ptr() if off is not None else None
        len_ptr, _keep = self._lengths_ptr(lengths, b, dev)
        io.lengths_dev = len_ptr
        io.own_length = int(bool(own_length) or off is not None)
        io.min_row_samples = int(min_samples) if min_samples is not None else n
        if row_gain is not None:
            _require_stat(row_gain, "row_gain", dev, b)
            io.row_gain_dev = row_gain.data_ptr()
        out = {}
        if want_codes:
            _require_stat(lo, "lo", dev, self.n_mels)
            _require_stat(scale, "scale", dev, self.n_mels)
            io.lo_dev, io.scale_dev, io.n_bins = lo.data_ptr(), scale.data_ptr(), int(n_bins)
            out["codes"] = torch.empty((b, self.n_mels, t), dtype=torch.uint8, device=dev)
            io.codes_dev = out["codes"].data_ptr()
            if want_mel_hat:
                _require_stat(step, "step", dev, self.n_mels)
                io.step_dev = step.data_ptr()
                out["mel_hat"] = torch.empty((b, self.n_mels, t), dtype=torch.float32, device=dev)
                io.mel_hat_dev = out["mel_hat"].data_ptr()
        if want_logmel:
            if logmel_dtype not in (torch.float32, torch.bfloat16):
                raise ValueError(f"logmel_dtype must be torch.float32 or torch.bfloat16, got {logmel_dtype}")
            out["logmel"] = torch.empty((b, self.n_mels, t), dtype=logmel_dtype, device=dev)
            io.logmel_dev, io.logmel_is_bf16 = out["logmel"].data_ptr(), int(logmel_dtype == torch.bfloat16)
        io.mask_invalid = int(mask_invalid)
        if run_min is not None or run_max is not None:
            _require_stat(run_min, "run_min", dev, self.n_mels)
            _require_stat(run_max, "run_max", dev, self.n_mels)
            io.min_dev, io.max_dev = run_min.data_ptr(), run_max.data_ptr()
        _native.check(_native.load().dmel_run(self._handle, ctypes.byref(io), _stream_ptr(dev)))
        return out

    def peak_gain(self, wav: torch.Tensor, *, lengths: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
                  n_rows: Optional[int] = None, max_samples: Optional[int] = None, target: float = 0.95) -> torch.Tensor:
        """(B,) float32 gains ``target / max|x|`` over the valid samples of every row (``dmel_row_peak_gain_f32``)."""
        _require_cuda(wav, "audio")
        dev = wav.device
        if offsets is not None:
            flat = wav.reshape(-1).float().contiguous()
            off = offsets.to(device=dev, dtype=torch.int64).contiguous()
            rows, b, n, stride = flat, int(n_rows), int(max_samples), int(max_samples)
        else:
            rows = as_rows(wav)
            b, n = rows.shape
            stride, off = (rows.stride(0) if b > 1 else n), None
        gain = torch.empty((b,), dtype=torch.float32, device=dev)
        len_ptr, _keep = self._lengths_ptr(lengths, b, dev)
        with torch.cuda.device(dev):
            _native.check(_native.load().dmel_row_peak_gain_f32(
                rows.data_ptr(), b, n, stride, off.data_ptr() if off is not None else None, len_ptr, float(target),
                gain.data_ptr(), _stream_ptr(dev)))
        return gain

    # -- helpers --------------------------------------------------------------
    def _frames_or_raise(self, n_samples: int) -> int:
        pad = (self.n_fft - self.hop_length) // 2
        if n_samples <= pad:
            raise ValueError(
                f"reflect padding of {pad} needs more than {pad} samples per row, got {n_samples}")
        t = self.num_frames(n_samples)
        if t <= 0:
            raise ValueError(f"{n_samples} samples are shorter than one frame of {self.n_fft}")
        return t

    def _lengths_ptr(self, lengths: Optional[torch.Tensor], b: int, device: torch.device):
        """Returns (pointer-or-None, keepalive)."""
        if lengths is None:
            return None, None
        flat = lengths.reshape(-1)
        if flat.numel() != b:
            raise ValueError(f"lengths has {flat.numel()} entries for a batch of {b}")
        converted = flat.to(device=device, dtype=torch.int32).contiguous()
        if converted.data_ptr() != flat.data_ptr() or converted.dtype != flat.dtype:
            # a temporary of ours: the launch is asynchronous, so tell the caching allocator that the launching
            # stream still reads it (two launches of one plan on two streams each keep their own lengths alive)
            converted.record_stream(torch.cuda.current_stream(device))
        return converted.data_ptr(), converted


# ---------------------------------------------------------------------------
# tensor-level quantiser stages (no plan needed)
# ---------------------------------------------------------------------------
def quantize(mel: torch.Tensor, lo: torch.Tensor, scale: torch.Tensor, n_bins: int,
             n_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``n_valid``: valid frames per batch row (B integers); frames at or past it get code 0 and are not read."""
    _require_cuda(mel, "mel")
    if mel.ndim != 3:
        raise ValueError(f"expected (B, n_mels, T), got {tuple(mel.shape)}")
    x = mel.float().contiguous()
    b, m, t = x.shape
    _require_stat(lo, "lo", x.device, m)
    _require_stat(scale, "scale", x.device, m)
    nv = None
    if n_valid is not None:
        nv = n_valid.detach().reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
        if nv.numel() != b:
            raise ValueError(f"n_valid has {nv.numel()} entries for a batch of {b}")
        nv.record_stream(torch.cuda.current_stream(x.device))
    codes = torch.empty((b, m, t), dtype=torch.uint8, device=x.device)
    if x.numel():
        _native.check(_native.load().dmel_quantize_masked_u8(
            x.data_ptr(), b, m, t, nv.data_ptr() if nv is not None else None, lo.data_ptr(), scale.data_ptr(), int(n_bins),
            codes.data_ptr(), _stream_ptr(x.device)))
    return codes


def quantizer_derive(lo: torch.Tensor, hi: torch.Tensor, n_bins: int):
    """(scale, step, ready) of a CUDA quantiser's statistics in one launch: K / (hi - lo) or 0, (hi - lo) / K, and an
    int32 device flag that is 1 iff lo <= hi on every channel."""
    _require_cuda(lo, "lo")
    m = lo.numel()
    _require_stat(lo, "lo", lo.device, m)
    _require_stat(hi, "hi", lo.device, m)
    scale, step = torch.empty_like(lo), torch.empty_like(lo)
    ready = torch.empty((), dtype=torch.int32, device=lo.device)
    _native.check(_native.load().dmel_quantizer_derive_f32(
        lo.data_ptr(), hi.data_ptr(), m, int(n_bins), scale.data_ptr(), step.data_ptr(), ready.data_ptr(), _stream_ptr(lo.device)))
    return scale, step, ready


def dequantize(codes: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    _require_cuda(codes, "codes")
    if codes.ndim != 3 or codes.dtype != torch.uint8:
        raise ValueError(f"expected uint8 codes (B, n_mels, T), got {codes.dtype} {tuple(codes.shape)}")
    c = codes.contiguous()
    b, m, t = c.shape
    if table.ndim != 2 or table.shape[0] != m:
        raise ValueError(f"codes have {m} channels, the quantiser table is {tuple(table.shape)}")
    _require_stat(table, "table", c.device, table.numel())
    out = torch.empty((b, m, t), dtype=torch.float32, device=c.device)
    if c.numel():
        _native.check(_native.load().dmel_dequantize_f32(
            c.data_ptr(), b, m, t, table.data_ptr(), int(table.shape[1]), out.data_ptr(),
            _stream_ptr(c.device)))
    return out


def tensor_minmax(mel: torch.Tensor, n_valid: Optional[torch.Tensor], run_min: torch.Tensor,
                  run_max: torch.Tensor) -> None:
    _require_cuda(mel, "mel")
    x = mel.float().contiguous()
    b, m, t = x.shape
    _require_stat(run_min, "run_min", x.device, m)
    _require_stat(run_max, "run_max", x.device, m)
    nv = None
    if n_valid is not None:
        nv = n_valid.reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
        if nv.numel() != b:
            raise ValueError(f"n_valid has {nv.numel()} entries for a batch of {b}")
    if x.numel():
        _native.check(_native.load().dmel_tensor_minmax_f32(
            x.data_ptr(), b, m, t, nv.data_ptr() if nv is not None else None, run_min.data_ptr(),
            run_max.data_ptr(), _stream_ptr(x.device)))
