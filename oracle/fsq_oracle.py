"""CPU oracle of the learned quantiser's elementwise core: finite scalar quantisation (FSQ) to indices, and the
language model's `id_shift`.  TEST INFRASTRUCTURE ONLY (see oracle/dmel_oracle.py for the rules).

The reference calls ``GroupedResidualFSQ(dim, levels=[7, 5, 5], num_quantizers=1, groups=10)`` of the third-party
package ``vector_quantize_pytorch>=1.20.9`` (reference setup.py:23; call sites models/modules/dowmsample_fsq.py:39-44,
:95, :130-137; config/lm/lm_config.yaml:95-107).  That package is neither vendored in /root/reference nor installed
or installable here, so this file restates the published algorithm of its ``FSQ`` class (Mentzer et al. 2023,
"Finite Scalar Quantization: VQ-VAE Made Simple", appendix A.1, which the package follows line by line) for the
part that has no learned weights:

    bound(z)      = tanh(z + shift) * half_l - offset,  half_l = (L - 1)(1 + eps)/2, offset = 0.5 for even L else 0,
                    shift = atanh(offset / half_l), eps = 1e-3
    quantize(z)   = round(bound(z)) / (L // 2)                      (the "codes", each in [-1, 1])
    index(codes)  = sum_d (codes_d * (L_d // 2) + L_d // 2) * basis_d,   basis = cumprod([1, L_0, L_1, ...])

**Parity unpinned** against the package itself (absent); pinned by hand-computed known answers in
tests/test_fsq.py.  The learned ``project_in`` / ``project_out`` linears of the package's ResidualFSQ (70 -> 3 -> 70
per group at the reference's sizes) and the residual scaling (one quantiser: scale 1) are outside this path.

``id_shift`` (reference models/modules/lm_process_input.py:301-313): ids[..., g] += g * audio_codebook_size.
"""
from __future__ import annotations

from typing import Sequence

import torch

EPS = 1e-3


def _consts(levels: Sequence[int]):
    lv = torch.tensor(list(levels), dtype=torch.float32)
    half_l = (lv - 1) * (1 + EPS) / 2
    offset = torch.where(lv % 2 == 0, torch.tensor(0.5), torch.tensor(0.0))
    shift = torch.atanh(offset / half_l)
    half_width = torch.tensor([l // 2 for l in levels], dtype=torch.float32)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1]), dtype=torch.int64), dim=0)
    return half_l, offset, shift, half_width, basis


def fsq_quantize(z: torch.Tensor, levels: Sequence[int]) -> torch.Tensor:
    """(..., D) float32 -> codes (..., D) in [-1, 1] (FSQ.quantize: round(bound(z)) / half_width)."""
    half_l, offset, shift, half_width, _ = _consts(levels)
    bounded = torch.tanh(z.float() + shift) * half_l - offset
    return torch.round(bounded) / half_width


def fsq_codes_to_indices(codes: torch.Tensor, levels: Sequence[int]) -> torch.Tensor:
    """(..., D) codes -> (...) int64 index in [0, prod(levels))."""
    _, _, _, half_width, basis = _consts(levels)
    scaled = codes * half_width + half_width
    return (scaled * basis.float()).sum(dim=-1).round().to(torch.int64)


def fsq_indices_to_codes(indices: torch.Tensor, levels: Sequence[int]) -> torch.Tensor:
    _, _, _, half_width, basis = _consts(levels)
    lv = torch.tensor(list(levels), dtype=torch.int64)
    digits = (indices[..., None] // basis) % lv
    return (digits.float() - half_width) / half_width


def grouped_fsq_encode(zp: torch.Tensor, levels: Sequence[int]):
    """zp (B, T, G, D): the per-group latents after the package's project_in.  -> (codes (B, T, G, D),
    indices (B, G, T) int64), the layout ``DownsampleFiniteScalarQuantize.encode`` returns for one quantiser per
    group ("g b l r -> b (g r) l", reference dowmsample_fsq.py:132)."""
    codes = fsq_quantize(zp, levels)
    return codes, fsq_codes_to_indices(codes, levels).permute(0, 2, 1).contiguous()


def id_shift(audio_ids: torch.Tensor, codebook_size: int) -> torch.Tensor:
    """(T, G) or (B, T, G) ids -> ids + g * codebook_size (reference lm_process_input.py:301-313)."""
    g = audio_ids.shape[-1]
    return audio_ids + torch.arange(g, dtype=audio_ids.dtype) * codebook_size
