"""Round-2 GPU parity hardening: full-size configs[4] and configs[2] geometries against
the oracle, the benchmark's fused-forward mode against the oracle directly, the frozen quantiser spec, and the
robustness cases the advisor listed.  All calls go through the C ABI.
"""
import importlib
import os
import sys
import textwrap

import numpy as np
import pytest
import torch

from conftest import GOLDEN_GEOMETRY, ROOT, logmel_close, oracle_config
from oracle import dmel_oracle as O
from test_gpu_parity import EDGE_EPS, REL_TOL, _check_codes, _tokenizer, _transform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def d(native_lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dmel_codec_b200
    return dmel_codec_b200


@pytest.fixture(scope="module")
def quant_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "quantizer_golden.npz"))


def test_derived_quantities_in_one_launch_equal_the_spec(d):
    """dmel_quantizer_derive_f32: scale = K / (hi - lo) or 0, step = (hi - lo) / K (float32 divisions, bit for bit what
    the torch formulas of SURVEY.md Appendix B give) and the calibrated flag, incl. a degenerate channel, an unseen
    channel (+inf / -inf) and an inverted one"""
    from dmel_codec_b200 import plan as P
    g = torch.Generator().manual_seed(5)
    lo = torch.randn(131, generator=g) * 3 - 8
    hi = lo + torch.rand(131, generator=g) * 12 + 1e-3
    hi[7] = lo[7]                                   # degenerate: scale 0
    for k in (1, 16, 32, 255, 256):
        scale, step, ready = P.quantizer_derive(lo.cuda(), hi.cuda(), k)
        width = hi - lo
        assert torch.equal(scale.cpu(), torch.where(width > 0, torch.full_like(width, float(k)) / width, torch.zeros_like(width)))
        assert torch.equal(step.cpu(), width / float(k)) and int(ready) == 1
    lo2, hi2 = lo.clone(), hi.clone()
    lo2[3], hi2[3] = float("inf"), float("-inf")    # a channel that never saw a frame
    scale, step, ready = P.quantizer_derive(lo2.cuda(), hi2.cuda(), 16)
    assert int(ready) == 0 and scale[3].item() == 0.0
    q = d.DMelQuantizer(131, 16).cuda()
    q.set_stats(lo, hi)
    assert q.calibrated and torch.equal(q.scale().cpu(), torch.where(width > 0, torch.full_like(width, 16.0) / width, torch.zeros_like(width)))
    q.set_stats(lo2, hi2)
    assert not q.calibrated


def test_deferred_calibration_check_still_raises(d):
    """encode(check_after=True) queues the launch before it learns whether the statistics are usable (the job's pass 2
    after the all-reduce); an uncalibrated quantiser must raise all the same, and work once calibrated"""
    q = d.DMelQuantizer(8, 16).cuda()
    mel = torch.randn(2, 8, 50, device="cuda")
    with pytest.raises(RuntimeError):
        q.encode(mel, check_after=True)
    with pytest.raises(RuntimeError):
        q.encode(mel)
    q.update_stats(mel)
    assert torch.equal(q.encode(mel, check_after=True), q.encode(mel))


@pytest.mark.parametrize("shape", [(5, 7, 33), (3, 80, 626), (4, 16, 1), (2, 3, 10), (6, 128, 937)])
@pytest.mark.parametrize("aligned", [True, False])
def test_length_aware_quantiser_equals_quantise_then_mask(d, shape, aligned):
    """dmel_quantize_masked_u8: code 0 at and past each row's valid-frame count (0, negative and oversized counts
    included), the plain quantiser's codes before it; groups of four that cross a channel or batch row, and tensors
    off a 16-byte boundary (the scalar path)"""
    from dmel_codec_b200 import plan as P
    b, m, t = shape
    g = torch.Generator().manual_seed(b * 1000 + m * 10 + t)
    n = b * m * t
    flat = torch.empty(n + 3, device="cuda")
    flat.copy_(torch.randn(n + 3, generator=g) * 4 - 5)
    mel = (flat[:n] if aligned else flat[3:3 + n]).view(b, m, t)
    lo = (torch.randn(m, generator=g) - 9).cuda()
    hi = lo + torch.rand(m, generator=g).cuda() * 10 + 0.5
    scale = torch.full_like(lo, 16.0) / (hi - lo)
    nv = torch.randint(-1, t + 3, (b,), generator=g, dtype=torch.int32)
    nv[0] = 0
    if b > 1:
        nv[1] = t
    plain = P.quantize(mel, lo, scale, 16)
    want = plain * (torch.arange(t, device="cuda")[None, None, :] < nv.cuda().clamp(0, t)[:, None, None])
    got = P.quantize(mel, lo, scale, 16, n_valid=nv)
    assert torch.equal(got, want)
    assert torch.equal(P.quantize(mel, lo, scale, 16, n_valid=torch.full((b,), t)), plain)


# ---------------------------------------------------------------------------
# (d) the frozen quantiser spec: stand-alone kernels bit for bit, fused encode up to edge ambiguity
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1_16k_80", "cfg2_24k_128", "cfg5_44k_160", "yaml_24k_100", "short_window", "edges"])
def test_quantiser_kernels_reproduce_the_frozen_spec(d, golden, quant_golden, name):
    mel = torch.from_numpy(quant_golden["edges/mel"] if name == "edges" else golden[name + "/logmel"]).cuda()
    k = int(quant_golden[name + "/n_bins"])
    q = d.DMelQuantizer(mel.shape[1], k).cuda()
    if name != "edges":  # calibration = exact min / max of the tensor
        q.update_stats(mel)
        assert torch.equal(q.lo.cpu(), torch.from_numpy(quant_golden[name + "/lo"]))
        assert torch.equal(q.hi.cpu(), torch.from_numpy(quant_golden[name + "/hi"]))
    q.set_stats(torch.from_numpy(quant_golden[name + "/lo"]), torch.from_numpy(quant_golden[name + "/hi"]))
    assert torch.equal(q.scale().cpu(), torch.from_numpy(quant_golden[name + "/scale"]))
    assert torch.equal(q.step().cpu(), torch.from_numpy(quant_golden[name + "/step"]))
    assert torch.equal(q.table().cpu(), torch.from_numpy(quant_golden[name + "/table"]))
    codes = q.encode(mel)
    assert torch.equal(codes.cpu(), torch.from_numpy(quant_golden[name + "/codes"]))
    assert torch.equal(q.decode(codes).cpu(), torch.from_numpy(quant_golden[name + "/decoded"]))


@pytest.mark.parametrize("name", ["cfg1_16k_80", "cfg2_24k_128", "cfg5_44k_160"])
def test_fused_encode_against_the_frozen_codes(d, golden, quant_golden, name):
    """waveform -> codes in one launch, with the frozen statistics: equal to the frozen codes (computed by numpy from
    the REFERENCE's log-mel) except where the reference's value sits within EDGE_EPS of an interior bin edge."""
    k = int(quant_golden[name + "/n_bins"])
    lo, hi = torch.from_numpy(quant_golden[name + "/lo"]), torch.from_numpy(quant_golden[name + "/hi"])
    tok = _tokenizer(d, GOLDEN_GEOMETRY[name], k)
    tok.quantizer.set_stats(lo, hi)
    codes, _ = tok.encode(torch.from_numpy(golden[name + "/wav"]).cuda())
    want = torch.from_numpy(quant_golden[name + "/codes"])
    mel_ref = torch.from_numpy(golden[name + "/logmel"])
    bad = codes.cpu() != want
    near = O.interior_edge_distance(mel_ref, lo, hi, k) < EDGE_EPS
    assert not torch.any(bad & ~near), f"{int((bad & ~near).sum())} mismatches away from any bin edge"


# ---------------------------------------------------------------------------
# (c) the benchmark's own mode (codes + dequantised mel in one launch) against the oracle, directly
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name,n_bins", [("cfg2_24k_128", 16), ("cfg5_44k_160", 32), ("cfg1_16k_80", 16)])
def test_fused_forward_against_the_oracle(d, golden, name, n_bins):
    kw = GOLDEN_GEOMETRY[name]
    wav = torch.from_numpy(golden[name + "/wav"])
    mel_ref = torch.from_numpy(golden[name + "/logmel"])
    lo, hi = O.calibrate_minmax(mel_ref)
    tok = _tokenizer(d, kw, n_bins)
    tok.quantizer.set_stats(lo, hi)
    res = tok.encode_decode(wav.cuda())
    _check_codes(res.codes, mel_ref, lo, hi, n_bins)
    # the dequantised mel is the oracle's decode of the codes the kernel emitted, bit for bit ...
    assert torch.equal(res.z.cpu(), O.dmel_decode(res.codes.cpu(), lo, hi, n_bins))
    # ... and within half a bin (+ the log-mel tolerance) of the reference's log-mel
    half = ((hi - lo) / n_bins / 2)[None, :, None]
    assert torch.all((res.z.cpu() - mel_ref).abs() <= half + 1e-4 * mel_ref.abs().clamp(min=1.0))


# ---------------------------------------------------------------------------
# (a) full-size configs[4]: one GPU's share, 4 x 60 s at 44.1 kHz, n_fft 2048 / hop 512 / 160 mel / 32 bins
# ---------------------------------------------------------------------------
def test_full_size_config5_share_against_the_oracle(d):
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg5_44k_160"]
    b, n = 4, 44100 * 60
    wav = synth.device_batch(range(900, 900 + b), n, 44100, "cuda")
    tok = _tokenizer(d, kw, 32)
    tok.calibrate([wav])
    codes, _, mel = tok.encode(wav, return_mel=True)
    assert codes.shape == (b, 160, 5167) and mel.shape == codes.shape
    lo, hi = tok.quantizer.lo.cpu(), tok.quantizer.hi.cpu()
    cfg = oracle_config(kw)
    # two whole rows (646 tiles each, the even/odd-split 2048 variant) against the oracle
    for row in (0, 3):
        ref = O.log_mel(wav[row:row + 1].cpu(), cfg)
        ok, ratio = logmel_close(mel[row:row + 1].cpu(), ref, REL_TOL)
        assert ok, f"row {row}: {ratio:.2f}x tolerance"
        _check_codes(codes[row:row + 1], ref, lo, hi, 32)
    # head and tail tiles of EVERY row: the reflected ends are where rows differ from the interior
    ref_all = O.log_mel(wav.cpu(), cfg)
    for sl in (slice(0, 24), slice(5167 - 24, 5167)):
        ok, ratio = logmel_close(mel[:, :, sl].cpu(), ref_all[:, :, sl], REL_TOL)
        assert ok, f"frames {sl}: {ratio:.2f}x tolerance"
    # fused forward at this size equals encode + decode
    res = tok.encode_decode(wav)
    assert torch.equal(res.codes, codes) and torch.equal(res.z, tok.decode(codes))


# ---------------------------------------------------------------------------
# (b) configs[2] geometry at scale: 256 utterances with random lengths, both job shapes against the oracle
# ---------------------------------------------------------------------------
def test_config3_geometry_with_random_lengths_against_the_oracle(d):
    from dmel_codec_b200 import distributed as D, synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    n_utts, n, bsz = 256, 16000 * 4, 64
    g = torch.Generator().manual_seed(123)
    lengths = torch.randint(16000, n + 1, (n_utts,), generator=g, dtype=torch.int32)
    lengths[5], lengths[77] = n, 385  # a full row and a row barely longer than the reflect pad
    wav = synth.device_batch(range(2000, 2000 + n_utts), n, 16000, "cuda")
    t = torch.arange(n, device="cuda")[None, None, :]
    wav = wav * (t < lengths.cuda()[:, None, None])  # right zero-padded rows, as the reference collate yields them
    lens_dev = lengths.cuda()
    load = lambda ids: (wav[ids[0]:ids[-1] + 1], lens_dev[ids[0]:ids[-1] + 1])
    cfg = oracle_config(kw)
    mel_ref = O.log_mel(wav.cpu(), cfg)
    n_valid = O.valid_frames(lengths.long(), 256)
    lo_ref, hi_ref = O.calibrate_minmax(mel_ref, n_valid)

    fused, single = _tokenizer(d, kw, 16), _tokenizer(d, kw, 16)
    D.calibrate_sharded(fused, n_utts, load, bsz)
    assert logmel_close(fused.quantizer.lo.cpu(), lo_ref, REL_TOL)[0] and logmel_close(fused.quantizer.hi.cpu(), hi_ref, REL_TOL)[0]
    assert fused.quantizer.lo.min().item() == pytest.approx(-11.512925148010254, abs=2e-6)  # the silent span of an utterance
    got_single = list(D.calibrate_encode_sharded(single, n_utts, load, bsz))
    assert torch.equal(fused.quantizer.lo, single.quantizer.lo) and torch.equal(fused.quantizer.hi, single.quantizer.hi)
    # codes with the ORACLE's statistics, so they compare value for value
    fused.quantizer.set_stats(lo_ref, hi_ref)
    for (ids, codes, code_lengths), (ids_s, codes_s, len_s) in zip(D.encode_sharded(fused, n_utts, load, bsz), got_single):
        rows = slice(ids[0], ids[-1] + 1)
        assert list(ids) == list(ids_s) and torch.equal(code_lengths.cpu().long(), n_valid[rows])
        _check_codes(codes, mel_ref[rows], lo_ref, hi_ref, 16, n_valid[rows])
        # the single-transform job quantised with ITS statistics: compare through the oracle with those
        _check_codes(codes_s, mel_ref[rows], single.quantizer.lo.cpu(), single.quantizer.hi.cpu(), 16, n_valid[rows])


# ---------------------------------------------------------------------------
# robustness
# ---------------------------------------------------------------------------
def test_two_live_plans_sharing_one_kernel_instantiation(d, golden):
    """Same n_fft, different n_mels and hop: the same template instantiation with different dynamic shared-memory
    sizes.  Alternating them used to leave the function's opt-in limit at the smaller plan's size."""
    kw_a = dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=128, f_min=0, f_max=12000)
    kw_b = dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=128, n_mels=40)
    wav = torch.from_numpy(golden["cfg2_24k_128/wav"]).cuda()
    a, b = _transform(d, kw_a), _transform(d, kw_b)
    dev = wav.device
    assert a.spectrogram.plan_for(dev).describe()["smem_bytes"] != b.spectrogram.plan_for(dev).describe()["smem_bytes"]
    first_a, first_b = a(wav), b(wav)
    for _ in range(3):
        assert torch.equal(a(wav), first_a)
        assert torch.equal(b(wav), first_b)
    torch.cuda.synchronize()
    ok, ratio = logmel_close(first_a.cpu(), torch.from_numpy(golden["cfg2_24k_128/logmel"]), REL_TOL)
    assert ok, ratio
    ok, ratio = logmel_close(first_b.cpu(), O.log_mel(wav.cpu(), oracle_config(kw_b)), REL_TOL)
    assert ok, ratio


def test_statistics_tensors_are_validated_before_the_launch(d, golden):
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    wav = torch.from_numpy(golden["cfg1_16k_80/wav"]).cuda()
    cpu_tok = d.DMelTokenizer(n_bins=16, **kw)  # buffers on the CPU: .cuda() forgotten
    with pytest.raises(ValueError, match="CUDA tensor"):
        cpu_tok.update_stats(wav)
    tok = _tokenizer(d, kw, 16)
    plan = tok._plan(wav.device)
    lo = torch.zeros(80, device="cuda")
    with pytest.raises(ValueError, match="float32"):
        plan.encode(wav, None, lo.double(), lo, 16)
    with pytest.raises(ValueError, match="at least 80"):
        plan.encode(wav, None, lo[:40], lo, 16)
    with pytest.raises(ValueError, match="contiguous"):
        plan.encode(wav, None, torch.zeros(160, device="cuda")[::2], lo, 16)
    with pytest.raises(ValueError, match="out must be"):
        tok.update_stats(wav)
        tok.encode_host(wav.cpu(), out=torch.empty(3, 3, dtype=torch.uint8))
    torch.cuda.synchronize()  # the context is still healthy
    assert torch.isfinite(tok.mel_transform(wav)).all()


def test_dtype_casts_of_a_parent_module_leave_the_statistics_in_float32(d, golden):
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    wav = torch.from_numpy(golden["cfg1_16k_80/wav"]).cuda()
    tok = _tokenizer(d, kw, 16)
    tok.update_stats(wav)
    want, _ = tok.encode(wav)
    lo = tok.quantizer.lo.clone()
    for cast in ("half", "bfloat16", "double"):
        getattr(tok, cast)()
        assert tok.quantizer.lo.dtype == torch.float32 and tok.quantizer.hi.dtype == torch.float32
    assert torch.equal(tok.quantizer.lo, lo)
    assert torch.equal(tok.encode(wav)[0], want)


def test_minus_zero_and_nan_do_not_corrupt_the_statistics(d):
    """atomic min/max on float bit patterns: -0.0 must not beat a negative minimum, NaN must not stick."""
    q = d.DMelQuantizer(4, 16).cuda()
    mel = torch.tensor([[[-5.0, -0.0, 3.0], [-0.0, 0.0, 0.0], [2.0, float("nan"), -1.0], [-0.0, -2.0, -7.0]]], device="cuda")
    q.update_stats(mel)
    assert q.lo.tolist() == [-5.0, 0.0, -1.0, -7.0] and q.hi.tolist()[0] == 3.0 and q.hi.tolist()[2] == 2.0
    assert q.hi[1].item() == 0.0 and q.hi[3].item() == 0.0  # (+0.0 == -0.0)


def test_plan_creation_failure_paths_do_not_leak_or_wedge(d):
    from dmel_codec_b200 import _native
    with pytest.raises(NotImplementedError):  # no kernel for this n_fft
        d.LogMelSpectrogram(sample_rate=16000, n_fft=4096, win_length=4096, hop_length=1024, n_mels=80)(torch.zeros(1, 16384, device="cuda"))
    with pytest.raises((ValueError, NotImplementedError, _native.DmelNativeError)):  # absurd channel count: no variant fits
        d.LogMelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=4000)(torch.zeros(1, 8192, device="cuda"))
    for _ in range(50):  # failed (or oversized) creations leave nothing behind that a later plan trips over
        try:
            d.LogMelSpectrogram(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=2048, n_mels=1024)(torch.zeros(1, 65536, device="cuda"))
        except (ValueError, NotImplementedError, _native.DmelNativeError):
            pass
    out = d.LogMelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80)(torch.zeros(1, 8192, device="cuda"))
    assert torch.isfinite(out).all()


def test_hbm_fit_fallback_of_the_single_transform_job(d, monkeypatch):
    """When the shard's log-mel does not fit, calibrate_encode_sharded runs as two transform passes: same result."""
    from dmel_codec_b200 import distributed as D, synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    pool = synth.batch(range(40, 46), 20000, 16000, "speech").cuda()
    load = lambda ids: pool[list(ids)]
    a, b = _tokenizer(d, kw, 16), _tokenizer(d, kw, 16)
    want = list(D.calibrate_encode_sharded(a, 6, load, 4))
    monkeypatch.setattr(D, "KEEP_MEL_HBM_FRACTION", 0.0)  # nothing fits
    got = list(D.calibrate_encode_sharded(b, 6, load, 4))
    assert torch.equal(a.quantizer.lo, b.quantizer.lo) and torch.equal(a.quantizer.hi, b.quantizer.hi)
    for (_, cw, _), (_, cg, _) in zip(want, got):
        assert torch.equal(cw, cg)
    # ... and when the estimate said yes but the allocation fails (another process took the memory in between)
    monkeypatch.setattr(D, "KEEP_MEL_HBM_FRACTION", 0.6)

    def no_memory(shape, device):
        raise torch.cuda.OutOfMemoryError("simulated")

    monkeypatch.setattr(D, "_alloc_store", no_memory)
    c = _tokenizer(d, kw, 16)
    again = list(D.calibrate_encode_sharded(c, 6, load, 4))
    assert torch.equal(a.quantizer.lo, c.quantizer.lo) and all(torch.equal(cw, cg) for (_, cw, _), (_, cg, _) in zip(want, again))


# ---------------------------------------------------------------------------
# the drop-in boundary: the reference's own `_target_` strings resolve to this implementation
# ---------------------------------------------------------------------------
def test_reference_config_targets_resolve_to_the_drop_in(d, golden, tmp_path, monkeypatch):
    """replace the body of dmel_codec/utils/spectrogram.py by one import line and every
    `_target_: dmel_codec.utils.spectrogram.LogMelSpectrogram` of the reference's configs (config/codec/dMel_used.yaml:88,98;
    config/lm/lm_config.yaml:115,125; config/lm/lm_inference.yaml:89,99) instantiates this implementation, with the
    keyword arguments those files pass (dMel_used.yaml:89-95)."""
    import yaml
    pkg = tmp_path / "dmel_codec" / "utils"
    pkg.mkdir(parents=True)
    (tmp_path / "dmel_codec" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("")
    (pkg / "spectrogram.py").write_text("from dmel_codec_b200.spectrogram import LinearSpectrogram, LogMelSpectrogram  # noqa: F401\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    for name in [m for m in sys.modules if m == "dmel_codec" or m.startswith("dmel_codec.")]:
        monkeypatch.delitem(sys.modules, name)
    node = yaml.safe_load(textwrap.dedent("""
        encode_mel_transform:
          _target_: dmel_codec.utils.spectrogram.LogMelSpectrogram
          sample_rate: 24000
          n_fft: 1024
          hop_length: 256
          win_length: 1024
          n_mels: 128
          f_min: 0
          f_max: 12000
    """))["encode_mel_transform"]
    module_name, _, cls_name = node.pop("_target_").rpartition(".")  # what hydra.utils.instantiate does with the string
    cls = getattr(importlib.import_module(module_name), cls_name)
    assert cls is d.LogMelSpectrogram
    transform = cls(**node)
    assert transform.hop_length == 256 and transform.sample_rate == 24000 and transform.n_mels == 128
    wav = torch.from_numpy(golden["cfg2_24k_128/wav"]).cuda()
    ok, ratio = logmel_close(transform(wav).cpu(), torch.from_numpy(golden["cfg2_24k_128/logmel"]), REL_TOL)
    assert ok, ratio


# ---------------------------------------------------------------------------
# the data module's host-side steps on the GPU - per-utterance peak normalisation and ragged
# (unpadded) batches.  Contract: every utterance equals the reference run on it alone.
# ---------------------------------------------------------------------------
def _utterances(lens, first_id, sr=16000, kind="speech"):
    from dmel_codec_b200 import synth
    return [synth.utterance(first_id + i, n, sr, kind) * (0.2 + 0.1 * i) for i, n in enumerate(lens)]  # different peaks


@pytest.mark.parametrize("peak", [None, 0.95])
@pytest.mark.parametrize("layout", ["ragged", "padded"])
def test_utterances_encode_as_if_alone(d, layout, peak):
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    lens = [16000, 385, 256 * 9 - 1, 256 * 9 + 1, 8191, 12000, 4097, 700]  # incl. the shortest legal one and k*hop +- 1
    utts = _utterances(lens, 3000)
    cfg = oracle_config(kw)
    ref = O.log_mel_each(utts, cfg, peak=peak)
    t_max = max(lens) // 256
    mel_ref = torch.zeros(len(lens), 80, t_max)
    for i, m in enumerate(ref):
        assert m.shape[1] == lens[i] // 256
        mel_ref[i, :, :m.shape[1]] = m
    n_valid = torch.tensor([n // 256 for n in lens])
    lo, hi = O.calibrate_minmax(mel_ref, n_valid)
    tok = _tokenizer(d, kw, 16)
    tok.quantizer.set_stats(lo, hi)
    if layout == "ragged":
        codes, code_lengths, mel = tok.encode_utterances([u.cuda() for u in utts], peak_normalize=peak, return_mel=True)
    else:
        padded = torch.zeros(len(lens), 1, max(lens))
        for i, u in enumerate(utts):
            padded[i, 0, :lens[i]] = u
        padded[:, :, :] += 0.0
        junk = padded.clone()
        for i, n in enumerate(lens):
            junk[i, 0, n:] = 0.37  # what lies past an utterance must not matter
        codes, code_lengths, mel = tok.encode_utterances(junk.cuda(), torch.tensor([lens], dtype=torch.int32).cuda(),
                                                         peak_normalize=peak, return_mel=True)
    assert codes.shape == (len(lens), 80, t_max) and torch.equal(code_lengths.cpu().long(), n_valid)
    t = torch.arange(t_max)[None, None, :]
    keep = (t < n_valid[:, None, None]).expand_as(mel_ref)
    got = mel.cpu()
    assert torch.all(got[~keep] == 0) and torch.all(codes.cpu()[~keep] == 0)
    ok, ratio = logmel_close(got[keep], mel_ref[keep], REL_TOL)
    assert ok, f"{layout} peak={peak}: {ratio:.2f}x tolerance"
    _check_codes(codes, mel_ref, lo, hi, 16, n_valid)


def test_peak_gain_matches_the_reference_normalisation(d):
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    lens = [5000, 16000, 9999, 1234]
    utts = _utterances(lens, 3100, kind="noise")
    utts[3] = torch.zeros(1234)  # a silent utterance: librosa divides by 1, the gain is the target itself
    padded = torch.zeros(4, max(lens))
    for i, u in enumerate(utts):
        padded[i, :lens[i]] = u
    padded[0, 5000:] = 0.9  # beyond the first utterance's length: must not count
    plan = _tokenizer(d, kw, 16)._plan(torch.device("cuda", torch.cuda.current_device()))
    gain = plan.peak_gain(padded.cuda(), lengths=torch.tensor(lens, dtype=torch.int32).cuda(), target=0.95).cpu()
    for i, u in enumerate(utts):
        peak = u.abs().max().item()
        want = 0.95 / peak if peak > 0 else 0.95
        assert gain[i].item() == pytest.approx(want, rel=2e-7), (i, gain[i].item(), want)
        # and the normalised samples agree with the reference's two-step arithmetic to a float32 ulp or two
        ours, ref = u * gain[i], O.peak_normalize(u)
        assert torch.all((ours - ref).abs() <= 2.4e-7 * ref.abs().clamp(min=1e-30))


def test_ragged_batch_rejects_an_utterance_shorter_than_the_reflect_pad(d):
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    tok = _tokenizer(d, kw, 16)
    tok.quantizer.set_stats(torch.full((80,), -11.0), torch.full((80,), 2.0))
    with pytest.raises(ValueError, match="reflect"):
        tok.encode_utterances([torch.zeros(4000).cuda(), torch.zeros(384).cuda()])


def test_zero_copy_streaming_equals_offline(d):
    """input_view / commit (the producer writes straight into the history buffer, one launch per chunk) emits the
    same codes as the offline encode of the whole waveform, across buffer compactions."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    n = 16000 * 6 + 77
    wav = synth.batch([31, 32], n, 16000, "speech").cuda()
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav])
    want, _ = tok.encode(wav)
    enc = d.DMelStreamEncoder(tok, n_streams=2, capacity_samples=8192)  # small: forces compaction every few chunks
    got, pos = [], 0
    for size in [1280] * 40 + [700, 3000, 333]:
        size = min(size, n - pos)
        if size <= 0:
            break
        view = enc.input_view(size)
        assert view.shape == (2, size) and view.is_cuda
        view.copy_(wav[:, 0, pos:pos + size])
        got.append(enc.commit(size))
        pos += size
    while pos < n:  # the rest through the copying push: both forms share the state
        size = min(1280, n - pos)
        got.append(enc.push(wav[:, 0, pos:pos + size]))
        pos += size
    got.append(enc.flush())
    assert torch.equal(torch.cat(got, dim=2), want)


# ---------------------------------------------------------------------------
# n_fft below 1024: the short frame runs zero-extended on the 1024-point kernel (exact bin subsampling)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [
    dict(sample_rate=16000, n_fft=512, win_length=512, hop_length=128, n_mels=80),
    dict(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=64),    # the classic 25 ms / 10 ms speech front end
    dict(sample_rate=8000, n_fft=256, win_length=256, hop_length=64, n_mels=40),
    dict(sample_rate=22050, n_fft=512, win_length=512, hop_length=512, n_mels=80),    # no overlap
    dict(sample_rate=16000, n_fft=128, win_length=128, hop_length=32, n_mels=20),
    dict(sample_rate=16000, n_fft=512, win_length=512, hop_length=128, n_mels=80, center=True),
], ids=["512_128", "512_win400_hop160", "256_64", "512_hop512", "128_32", "512_center"])
def test_small_n_fft_against_the_oracle(d, kw):
    from dmel_codec_b200 import synth
    wav = synth.batch(range(500, 503), 20011, kw["sample_rate"], "speech")
    ref = O.log_mel(wav, oracle_config(kw))
    tr = _transform(d, kw)
    got = tr(wav.cuda()).cpu()
    assert got.shape == ref.shape
    ok, ratio = logmel_close(got, ref, REL_TOL)
    assert ok, f"{ratio:.2f}x tolerance"
    info = tr.spectrogram.plan_for(torch.device("cuda", torch.cuda.current_device())).describe()
    assert info["n_fft"] == kw["n_fft"] and info["core_fft"] == 1024
    # and the quantiser path on top of it
    lo, hi = O.calibrate_minmax(ref)
    tok = _tokenizer(d, kw, 16)
    tok.quantizer.set_stats(lo, hi)
    codes, _ = tok.encode(wav.cuda())
    _check_codes(codes, ref, lo, hi, 16)
    # streaming over the short transform: same codes as offline
    if not kw.get("center"):
        enc = d.DMelStreamEncoder(tok, n_streams=3, capacity_samples=8192)
        outs, pos = [], 0
        while pos < wav.shape[2]:
            outs.append(enc.push(wav[:, 0, pos:pos + 1000].cuda()))
            pos += 1000
        outs.append(enc.flush())
        assert torch.equal(torch.cat(outs, dim=2), codes)


def test_unsupported_n_fft_is_refused(d):
    for n_fft in (400, 1000, 4096, 32):
        with pytest.raises(NotImplementedError, match="n_fft"):
            d.LogMelSpectrogram(sample_rate=16000, n_fft=n_fft, win_length=min(n_fft, 400), hop_length=100, n_mels=40)(
                torch.zeros(1, 8000, device="cuda"))
