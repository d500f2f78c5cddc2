"""CPU tests: the oracle against the reference's golden vectors, the product's
host-side constants against the oracle, and the quantiser spec's properties."""
import numpy as np
import pytest
import torch

import os

from conftest import GOLDEN_GEOMETRY, ROOT, oracle_config
from dmel_codec_b200 import filters, synth
from oracle import dmel_oracle as O


@pytest.mark.parametrize("name", sorted(GOLDEN_GEOMETRY))
def test_oracle_matches_reference_golden(golden, name):
    cfg = oracle_config(GOLDEN_GEOMETRY[name])
    wav = torch.from_numpy(golden[name + "/wav"])
    ref = torch.from_numpy(golden[name + "/logmel"])
    got = O.log_mel(wav, cfg)
    assert got.shape == ref.shape
    assert got.shape[2] == cfg.n_frames(wav.shape[-1])
    # same torch build, same ops: equal to the last bit here; allow 1e-6 for another BLAS/FFT build
    assert (got - ref).abs().max().item() <= 1e-6


def test_silence_hits_exact_floor(golden):
    ref = golden["silence/logmel"]
    assert np.all(ref == np.float32(np.log(np.float32(1e-5))))
    assert float(ref.flat[0]) == pytest.approx(-11.512925148010254, abs=1e-6)


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [
    (16000, 1024, 80, 0.0, None), (24000, 1024, 128, 0.0, 12000.0), (44100, 2048, 160, 0.0, None),
    (24000, 1024, 100, 0.0, 12000.0), (22050, 1024, 64, 30.0, 8000.0)])
def test_filterbank_restatements_agree(sr, n_fft, n_mels, fmin, fmax):
    a = O.slaney_filterbank(sr, n_fft, n_mels, fmin, fmax)
    b = filters.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    assert a.dtype == b.dtype == np.float32 and a.shape == (n_mels, n_fft // 2 + 1)
    assert np.array_equal(a, b)
    # independent implementation of the same published definition
    import torchaudio.functional as AF
    t = AF.melscale_fbanks(n_fft // 2 + 1, fmin, fmax if fmax else sr / 2, n_mels, sr, norm="slaney",
                           mel_scale="slaney").T.numpy()
    assert np.abs(a - t).max() < 5e-7


def test_window_matches_torch_and_centres_short_windows():
    assert np.array_equal(filters.stft_window(1024, 1024), torch.hann_window(1024).numpy())
    w = filters.stft_window(800, 1024)
    assert w.shape == (1024,) and np.all(w[:112] == 0) and np.all(w[912:] == 0)
    assert np.array_equal(w[112:912], torch.hann_window(800).numpy())
    assert np.array_equal(w, O.stft_window(800, 1024).numpy())


def test_banded_bank_reconstructs_dense():
    bank = filters.mel_filterbank(24000, 1024, 128, 0.0, 12000.0)
    band = filters.band_filterbank(bank)
    dense = np.zeros_like(bank)
    for m in range(bank.shape[0]):
        s, c, o = band.start[m], band.count[m], band.offset[m]
        assert c % 4 == 0 and o % 4 == 0
        n = min(c, bank.shape[1] - s)
        dense[m, s:s + n] = band.weights[o:o + n]
        assert np.all(band.weights[o + n:o + c] == 0)
    assert np.array_equal(dense, bank)


def test_reflect_index_matches_torch_pad():
    x = torch.arange(10.0)
    padded = torch.nn.functional.pad(x[None, None], (4, 4), mode="reflect")[0, 0]
    idx = O.reflect_index(np.arange(-4, 14), 10)
    assert np.array_equal(padded.numpy(), x.numpy()[idx])


def test_frame_count_formula():
    cfg = oracle_config(GOLDEN_GEOMETRY["cfg1_16k_80"])
    for n in (160000, 160001, 159999, 1280, 1024, 500, 385):
        assert cfg.n_frames(n) == n // 256
    with pytest.raises(ValueError):
        O.log_mel(torch.zeros(1, 384), cfg)


# ---- quantiser spec ---------------------------------
def _mel_and_stats(n_bins=16):
    cfg = oracle_config(GOLDEN_GEOMETRY["cfg1_16k_80"])
    wav = synth.batch(range(3), 16000, 16000)
    mel = O.log_mel(wav, cfg)
    lo, hi = O.calibrate_minmax(mel)
    return mel, lo, hi


@pytest.mark.parametrize("n_bins", [2, 16, 32, 256])
def test_codes_in_range_and_roundtrip_within_half_bin(n_bins):
    mel, lo, hi = _mel_and_stats()
    codes = O.dmel_encode(mel, lo, hi, n_bins)
    assert codes.dtype == torch.uint8 and codes.max().item() <= n_bins - 1
    back = O.dmel_decode(codes, lo, hi, n_bins)
    half = ((hi - lo) / n_bins / 2)[None, :, None]
    assert torch.all((back - mel).abs() <= half * (1 + 1e-4) + 1e-6)
    # min maps to bin 0, max to the last bin (the clamp)
    assert codes.amin(dim=(0, 2)).eq(0).all() and codes.amax(dim=(0, 2)).eq(n_bins - 1).all()


def test_encode_is_monotone_per_channel():
    lo, hi = torch.tensor([-11.5, -3.0]), torch.tensor([1.0, 2.0])
    x = torch.linspace(-12, 3, 4001)[None, None, :].expand(1, 2, -1).contiguous()
    codes = O.dmel_encode(x, lo, hi, 16).long()
    assert torch.all(codes[..., 1:] >= codes[..., :-1])


def test_masked_calibration_ignores_padding():
    mel, _, _ = _mel_and_stats()
    n_valid = torch.tensor([10, 62, 0])
    lo, hi = O.calibrate_minmax(mel, n_valid)
    keep = torch.cat([mel[0, :, :10], mel[1, :, :62]], dim=1)
    assert torch.equal(lo, keep.amin(dim=1)) and torch.equal(hi, keep.amax(dim=1))


def test_degenerate_channel_encodes_to_zero():
    mel = torch.full((1, 2, 5), -11.5)
    lo, hi = O.calibrate_minmax(mel)
    assert O.dmel_encode(mel, lo, hi, 16).eq(0).all()
    assert torch.equal(O.dmel_decode(torch.zeros(1, 2, 5, dtype=torch.uint8), lo, hi, 16), mel)


def test_synth_is_seeded_and_bounded():
    a, b = synth.utterance(5, 4000, 16000), synth.utterance(5, 4000, 16000)
    assert torch.equal(a, b) and a.abs().max() <= 0.95 + 1e-6
    assert not torch.equal(a, synth.utterance(6, 4000, 16000))
    assert (a == 0).sum() >= 4000 // 20  # the silent span


# ---------------------------------------------------------------------------
# the quantiser spec is frozen: tests/golden/quantizer_golden.npz (plain numpy, make_quantizer_golden.py)
# ---------------------------------------------------------------------------
QUANT_CASES = ["cfg1_16k_80", "cfg2_24k_128", "cfg5_44k_160", "yaml_24k_100", "short_window"]


@pytest.fixture(scope="module")
def quant_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "quantizer_golden.npz"))


@pytest.mark.parametrize("name", QUANT_CASES)
def test_oracle_quantiser_reproduces_the_frozen_spec(golden, quant_golden, name):
    mel = torch.from_numpy(golden[name + "/logmel"])
    k = int(quant_golden[name + "/n_bins"])
    lo, hi = O.calibrate_minmax(mel)
    assert torch.equal(lo, torch.from_numpy(quant_golden[name + "/lo"])) and torch.equal(hi, torch.from_numpy(quant_golden[name + "/hi"]))
    assert torch.equal(O.bin_scale(lo, hi, k), torch.from_numpy(quant_golden[name + "/scale"]))
    codes = O.dmel_encode(mel, lo, hi, k)
    assert torch.equal(codes, torch.from_numpy(quant_golden[name + "/codes"]))
    assert torch.equal(O.dmel_decode_table(lo, hi, k), torch.from_numpy(quant_golden[name + "/table"]))
    assert torch.equal(O.dmel_decode(codes, lo, hi, k), torch.from_numpy(quant_golden[name + "/decoded"]))


def test_oracle_quantiser_edge_cases_match_the_frozen_spec(quant_golden):
    """degenerate channel, value == hi / lo, values on and one ulp around every interior edge, out-of-range values"""
    mel = torch.from_numpy(quant_golden["edges/mel"])
    lo, hi = torch.from_numpy(quant_golden["edges/lo"]), torch.from_numpy(quant_golden["edges/hi"])
    codes = O.dmel_encode(mel, lo, hi, 16)
    assert torch.equal(codes, torch.from_numpy(quant_golden["edges/codes"]))
    assert torch.equal(O.dmel_decode(codes, lo, hi, 16), torch.from_numpy(quant_golden["edges/decoded"]))
    assert torch.all(codes[0, 0] == 0)                 # degenerate channel: scale 0, every code 0
    assert torch.all(codes[0, 1:, 1] == 15)            # x == hi lands in the last bin
    assert torch.all(codes[0, 1:, 0] == 0) and torch.all(codes[0, 1:, 2] == 0) and torch.all(codes[0, 1:, 3] == 15)
