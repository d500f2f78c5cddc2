"""Finite-scalar-quantisation indices on the GPU: the weight-free core of the reference's
learned quantiser, ``GroupedResidualFSQ(levels=[7, 5, 5], num_quantizers=1, groups=10)`` of
``vector_quantize_pytorch`` (reference models/modules/dowmsample_fsq.py:39-44; encode at :124-133, decode at :135-147),
with the language model's ``id_shift`` (reference models/modules/lm_process_input.py:301-313) fused in.

The learned ``project_in`` / ``project_out`` linears that surround this step stay with the reference's modules; what
runs here is what lies between them: latents (B, T, G, D) -> codes in [-1, 1] and integer indices, and back.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _native


class FSQIndexer:
    def __init__(self, levels: Sequence[int] = (7, 5, 5), groups: int = 10):
        self.levels = tuple(int(l) for l in levels)
        self.groups = int(groups)
        self._levels_c = (ctypes.c_int * len(self.levels))(*self.levels)
        self.codebook_size = 1
        for l in self.levels:
            self.codebook_size *= l

    def _check(self, zp: torch.Tensor) -> Tuple[int, int]:
        if not zp.is_cuda:
            raise RuntimeError("FSQIndexer runs on CUDA tensors only (no CPU fallback)")
        if zp.ndim != 4 or zp.shape[2] != self.groups or zp.shape[3] != len(self.levels):
            raise ValueError(f"expected latents (B, T, {self.groups}, {len(self.levels)}), got {tuple(zp.shape)}")
        return zp.shape[0], zp.shape[1]

    @torch.no_grad()
    def encode(self, zp: torch.Tensor, *, return_codes: bool = True, lm_codebook_size: Optional[int] = None):
        """zp (B, T, G, D) -> dict with ``indices`` (B, G, T) int64 (the layout the reference's ``encode`` returns),
        optionally ``codes`` (B, T, G, D) and ``lm_ids`` (B, T, G) = index + g * lm_codebook_size."""
        b, t = self._check(zp)
        z = zp.float().contiguous()
        out = {"indices": torch.empty((b, self.groups, t), dtype=torch.int64, device=z.device)}
        if return_codes:
            out["codes"] = torch.empty_like(z)
        if lm_codebook_size is not None:
            out["lm_ids"] = torch.empty((b, t, self.groups), dtype=torch.int64, device=z.device)
        if b and t:
            _native.check(_native.load().dmel_fsq_encode(
                z.data_ptr(), b, t, self.groups, self._levels_c, len(self.levels),
                out["codes"].data_ptr() if return_codes else None, out["indices"].data_ptr(),
                out["lm_ids"].data_ptr() if lm_codebook_size is not None else None,
                int(lm_codebook_size or 0), torch.cuda.current_stream(z.device).cuda_stream))
        return out

    @torch.no_grad()
    def decode(self, indices: torch.Tensor) -> torch.Tensor:
        """indices (B, G, T) int64 -> codes (B, T, G, D) float32 (``indices_to_codes``, before the learned project_out)."""
        if not indices.is_cuda or indices.ndim != 3 or indices.shape[1] != self.groups or indices.dtype != torch.int64:
            raise ValueError(f"expected CUDA int64 indices (B, {self.groups}, T), got {indices.dtype} {tuple(indices.shape)}")
        idx = indices.contiguous()
        b, _, t = idx.shape
        codes = torch.empty((b, t, self.groups, len(self.levels)), dtype=torch.float32, device=idx.device)
        if b and t:
            _native.check(_native.load().dmel_fsq_decode(idx.data_ptr(), b, t, self.groups, self._levels_c, len(self.levels),
                                                         codes.data_ptr(), torch.cuda.current_stream(idx.device).cuda_stream))
        return codes
