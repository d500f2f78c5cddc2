"""FSQ codes / indices / id_shift.  The algorithm lives in the un-vendored package
vector_quantize_pytorch (>= 1.20.9, reference setup.py:23), absent here: the oracle restates the published FSQ
arithmetic and is pinned by hand-computed known answers below; the CUDA kernels are compared with the oracle."""
import pytest
import torch

from oracle import fsq_oracle as F

LEVELS = (7, 5, 5)  # reference config/lm/lm_config.yaml:100-103


def test_oracle_known_answers():
    # z = 0: bound = 0 in every dimension (odd levels: no offset) -> digits (3, 2, 2) -> 3 + 2*7 + 2*35 = 87
    assert F.fsq_codes_to_indices(F.fsq_quantize(torch.zeros(1, 3), LEVELS), LEVELS).item() == 87
    # saturated inputs reach the extreme digits: +inf-like -> (6, 4, 4) -> 6 + 28 + 140 = 174; -inf-like -> 0
    assert F.fsq_codes_to_indices(F.fsq_quantize(torch.full((1, 3), 20.0), LEVELS), LEVELS).item() == 174
    assert F.fsq_codes_to_indices(F.fsq_quantize(torch.full((1, 3), -20.0), LEVELS), LEVELS).item() == 0
    # codes are multiples of 1 / (L // 2): atanh(1/3.003) makes tanh(z) * 3.003 = 1 exactly on the 7-level axis
    z = torch.tensor([[torch.atanh(torch.tensor(1.0 / 3.003)).item(), 0.0, 0.0]])
    codes = F.fsq_quantize(z, LEVELS)
    assert torch.allclose(codes, torch.tensor([[1.0 / 3.0, 0.0, 0.0]]))
    assert F.fsq_codes_to_indices(codes, LEVELS).item() == 4 + 14 + 70
    # even levels carry the half-step offset: L = 8 -> digits 0..7, z = 0 -> bounded = tanh(shift)*3.5035 - 0.5 = 0 -> digit 4
    assert F.fsq_codes_to_indices(F.fsq_quantize(torch.zeros(1, 1), (8,)), (8,)).item() == 4


def test_oracle_round_trip_and_range():
    g = torch.Generator().manual_seed(0)
    z = torch.randn(4, 50, 10, 3, generator=g) * 2
    codes, idx = F.grouped_fsq_encode(z, LEVELS)
    assert idx.shape == (4, 10, 50) and idx.min() >= 0 and idx.max() < 175
    back = F.fsq_indices_to_codes(idx.permute(0, 2, 1), LEVELS)
    assert torch.allclose(back, codes)
    ids = F.id_shift(idx.permute(0, 2, 1), 180)
    assert torch.equal(ids[..., 3] - idx.permute(0, 2, 1)[..., 3], torch.full((4, 50), 540))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,levels", [((3, 201, 10, 3), (7, 5, 5)), ((1, 64, 1, 4), (8, 5, 5, 5)), ((2, 7, 16, 2), (4, 9)),
                                          ((2, 300, 3, 1), (16,)), ((1, 515, 7, 5), (3, 4, 5, 6, 7)), ((2, 33, 2, 6), (2, 3, 2, 3, 2, 3)),
                                          ((1, 40, 5, 7), (3, 3, 3, 3, 3, 3, 3)), ((2, 1000, 12, 8), (4, 3, 3, 3, 2, 2, 2, 2))])
def test_cuda_fsq_matches_oracle(native_lib, shape, levels):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dmel_codec_b200 as d
    g = torch.Generator().manual_seed(1)
    z = torch.randn(*shape, generator=g) * 1.5
    z[0, 0] = 0.0
    z[0, 1] = 30.0
    z[0, 2] = -30.0
    fsq = d.FSQIndexer(levels=levels, groups=shape[2])
    out = fsq.encode(z.cuda(), lm_codebook_size=180)
    codes_ref, idx_ref = F.grouped_fsq_encode(z, levels)
    # tanhf on the GPU and torch.tanh on the CPU may differ in the last bit: only a value within 1e-5 of a rounding
    # boundary (x.5) may land in the neighbouring level
    half_l, offset, shift, half_width, _ = F._consts(levels)
    bounded = torch.tanh(z + shift) * half_l - offset
    near = ((bounded - torch.floor(bounded) - 0.5).abs() < 1e-5).any(dim=-1)
    bad = (out["codes"].cpu() != codes_ref).any(dim=-1)
    assert not torch.any(bad & ~near)
    idx = out["indices"].cpu()
    assert torch.equal(idx.permute(0, 2, 1)[~near], idx_ref.permute(0, 2, 1)[~near])
    assert torch.equal(out["lm_ids"].cpu(), F.id_shift(idx.permute(0, 2, 1), 180))
    # indices <-> codes are exact inverses of each other on the GPU, and agree with the oracle's inverse
    assert torch.equal(fsq.decode(out["indices"]).cpu(), F.fsq_indices_to_codes(idx.permute(0, 2, 1), levels))
    assert torch.equal(fsq.decode(out["indices"]), out["codes"])
