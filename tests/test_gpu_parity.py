"""GPU parity tests: the CUDA path (through the C ABI) against the reference's
golden vectors and against the CPU oracle on the same seeded inputs.

Gates:
  log-mel   |a - b| <= 1e-4 * max(1, |b|)
  codes     bit exact, except values within EDGE_EPS of an interior bin edge
  decode    bit exact (table lookup of oracle-computed centres)
  stats     bit exact against torch.amin/amax over the kernel's own log-mel
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_GEOMETRY, logmel_close, oracle_config
from oracle import dmel_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
EDGE_EPS = 1e-4  # log-mel units


@pytest.fixture(scope="module")
def d(native_lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dmel_codec_b200
    return dmel_codec_b200


def _transform(d, kw):
    return d.LogMelSpectrogram(**kw)


# ---------------------------------------------------------------------------
# log-mel against the reference's own outputs
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GOLDEN_GEOMETRY))
def test_logmel_matches_reference_golden(d, golden, name):
    kw = GOLDEN_GEOMETRY[name]
    wav = torch.from_numpy(golden[name + "/wav"]).cuda()
    ref = torch.from_numpy(golden[name + "/logmel"])
    got = _transform(d, kw)(wav)
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == tuple(ref.shape)
    ok, ratio = logmel_close(got.cpu(), ref, REL_TOL)
    assert ok, f"{name}: worst error is {ratio:.2f}x the tolerance"


def test_silence_is_exactly_the_floor(d, golden):
    wav = torch.from_numpy(golden["silence/wav"]).cuda()
    got = _transform(d, GOLDEN_GEOMETRY["silence"])(wav)
    assert torch.all(got == got.flatten()[0])
    assert abs(got.flatten()[0].item() - (-11.512925148010254)) < 2e-6


def test_2d_and_3d_inputs_agree(d, golden):
    kw = GOLDEN_GEOMETRY["cfg2_24k_128"]
    wav = torch.from_numpy(golden["cfg2_24k_128/wav"]).cuda()
    tr = _transform(d, kw)
    assert torch.equal(tr(wav), tr(wav[:, 0, :]))
    # a strided view (rows of a wider buffer) takes the row_stride path
    wide = torch.zeros(wav.shape[0], wav.shape[2] + 37, device="cuda")
    wide[:, :wav.shape[2]] = wav[:, 0, :]
    assert torch.equal(tr(wide[:, :wav.shape[2]]), tr(wav))


@pytest.mark.parametrize("n_samples", [385, 511, 1024, 1279, 4097, 8192 + 255, 40000])
def test_logmel_vs_oracle_ragged_lengths(d, n_samples):
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    wav = synth.batch(range(100, 103), n_samples, 16000, "noise")
    ref = O.log_mel(wav, oracle_config(kw))
    got = _transform(d, kw)(wav.cuda()).cpu()
    assert got.shape == ref.shape
    ok, ratio = logmel_close(got, ref, REL_TOL)
    assert ok, f"L={n_samples}: {ratio:.2f}x tolerance"


@pytest.mark.parametrize("kw", [
    dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=128),   # reference defaults
    dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=441, n_mels=80),    # hop not dividing n_fft (odd pad)
    dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160, n_mels=80),
    dict(sample_rate=16000, n_fft=1024, win_length=400, hop_length=1024, n_mels=40),    # no overlap at all
    dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=128, center=True),
], ids=["ref_default", "hop441", "hop160", "hop_eq_nfft", "center"])
def test_logmel_vs_oracle_other_geometries(d, kw):
    from dmel_codec_b200 import synth
    wav = synth.batch(range(200, 204), 30011, kw["sample_rate"], "speech")
    ref = O.log_mel(wav, oracle_config(kw))
    got = _transform(d, kw)(wav.cuda()).cpu()
    assert got.shape == ref.shape
    ok, ratio = logmel_close(got, ref, REL_TOL)
    assert ok, f"{ratio:.2f}x tolerance"


def test_launch_configuration_for_the_benchmark_geometry(d):
    """configs[1] must get the register-lean 8-frame tile with three CTAs per SM (its shared memory is sized
    to fit a third of the SM), and the 2048 geometry the two-CTA even/odd-split variant."""
    dev = torch.device("cuda", torch.cuda.current_device())
    info = _transform(d, GOLDEN_GEOMETRY["cfg2_24k_128"]).spectrogram.plan_for(dev).describe()
    assert info["tile_frames"] == 8 and info["ctas_per_sm"] == 3, info
    assert info["smem_bytes"] <= 233472 // 3 - 1024
    info = _transform(d, GOLDEN_GEOMETRY["cfg5_44k_160"]).spectrogram.plan_for(dev).describe()
    assert info["tile_frames"] == 8 and info["ctas_per_sm"] == 2, info


@pytest.mark.parametrize("occ", ["1", "2"])
def test_fallback_kernel_variants_agree_with_golden(d, golden, occ, monkeypatch):
    """The variants chosen when shared memory is tight (fewer CTAs per SM, other FFT forms) stay correct."""
    monkeypatch.setenv("DMEL_OCC", occ)
    for name in ("cfg2_24k_128", "cfg5_44k_160", "short_window"):
        if occ == "2" and name == "cfg5_44k_160":
            continue  # that is already the default variant
        kw = GOLDEN_GEOMETRY[name]
        tr = _transform(d, kw)
        got = tr(torch.from_numpy(golden[name + "/wav"]).cuda())
        info = tr.spectrogram.plan_for(got.device).describe()
        assert info["ctas_per_sm"] == int(occ), info
        ok, ratio = logmel_close(got.cpu(), torch.from_numpy(golden[name + "/logmel"]), REL_TOL)
        assert ok, f"{name} occ={occ}: {ratio:.2f}x tolerance"


def test_too_short_input_raises_like_the_reference(d):
    tr = _transform(d, GOLDEN_GEOMETRY["cfg1_16k_80"])
    with pytest.raises(ValueError, match="reflect"):
        tr(torch.zeros(2, 384, device="cuda"))  # pad is 384: F.pad(reflect) rejects L <= pad
    with pytest.raises(ValueError):
        tr(torch.zeros(2, 2, 4096, device="cuda"))  # not mono


# ---------------------------------------------------------------------------
# calibration, codes, decode
# ---------------------------------------------------------------------------
def _tokenizer(d, kw, n_bins):
    args = dict(kw)
    return d.DMelTokenizer(n_bins=n_bins, **args).cuda()


def _check_codes(codes_gpu, mel_ref, lo, hi, n_bins, n_valid=None):
    """Every mismatch against the oracle's codes must sit within EDGE_EPS of an interior edge."""
    want = O.dmel_encode(mel_ref, lo, hi, n_bins)
    got = codes_gpu.cpu()
    if n_valid is not None:
        t = torch.arange(got.shape[2])[None, None, :]
        keep = (t < n_valid.reshape(-1, 1, 1)).expand_as(got)
        assert torch.all(got[~keep] == 0)
    else:
        keep = torch.ones_like(got, dtype=torch.bool)
    bad = (got != want) & keep
    near = O.interior_edge_distance(mel_ref, lo, hi, n_bins) < EDGE_EPS
    assert not torch.any(bad & ~near), f"{int((bad & ~near).sum())} code mismatches away from any bin edge"
    assert torch.all((got.int() - want.int()).abs()[bad] == 1)
    return int(bad.sum()), int((near & keep).sum()), int(keep.sum())


@pytest.mark.parametrize("name,n_bins", [("cfg1_16k_80", 16), ("cfg2_24k_128", 16), ("cfg5_44k_160", 32),
                                         ("yaml_24k_100", 16)])
def test_calibrate_encode_decode_against_oracle(d, golden, name, n_bins):
    kw = GOLDEN_GEOMETRY[name]
    wav = torch.from_numpy(golden[name + "/wav"])
    mel_ref = torch.from_numpy(golden[name + "/logmel"])
    tok = _tokenizer(d, kw, n_bins)

    # calibration: bit exact against amin/amax of the kernel's own log-mel; close to the oracle's
    tok.update_stats(wav.cuda())
    own = tok.mel_transform(wav.cuda())
    assert torch.equal(tok.quantizer.lo, own.amin(dim=(0, 2))) and torch.equal(tok.quantizer.hi, own.amax(dim=(0, 2)))
    lo_ref, hi_ref = O.calibrate_minmax(mel_ref)
    assert logmel_close(tok.quantizer.lo.cpu(), lo_ref, REL_TOL)[0] and logmel_close(tok.quantizer.hi.cpu(), hi_ref, REL_TOL)[0]

    # encode with the ORACLE's stats so codes are comparable value for value
    tok.quantizer.set_stats(lo_ref, hi_ref)
    near_edge = torch.zeros(1, dtype=torch.int64, device="cuda")
    codes, code_lengths, mel_gpu = tok.encode(wav.cuda(), return_mel=True, near_edge=near_edge, edge_eps=EDGE_EPS)
    assert codes.dtype == torch.uint8 and code_lengths is None
    assert torch.equal(mel_gpu, own)
    bad, near, total = _check_codes(codes, mel_ref, lo_ref, hi_ref, n_bins)
    assert bad <= near
    # the kernel's own near-edge counter agrees with a recount from its own log-mel
    recount = int((O.interior_edge_distance(mel_gpu.cpu(), lo_ref, hi_ref, n_bins) < EDGE_EPS).sum())
    assert abs(int(near_edge.item()) - recount) <= max(2, recount // 50)

    # fused encode == stand-alone quantiser on the kernel's own log-mel, bit for bit
    assert torch.equal(codes, tok.quantizer.encode(mel_gpu))
    # decode: bit exact against the oracle
    assert torch.equal(tok.decode(codes).cpu(), O.dmel_decode(codes.cpu(), lo_ref, hi_ref, n_bins))
    res = tok.quantizer(mel_gpu)
    assert torch.equal(res.codes, codes) and torch.equal(res.z, tok.decode(codes)) and res.latents is mel_gpu


def test_lengths_mask_calibration_and_codes(d):
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    lengths = torch.tensor([[20000, 7000, 256 * 30 + 5, 100]], dtype=torch.int32)  # (1, B) like the reference collate
    wav = synth.batch(range(300, 304), 20000, 16000, "speech", lengths=lengths[0].tolist())
    mel_ref = O.log_mel(wav, oracle_config(kw))
    n_valid = O.valid_frames(lengths[0].long(), 256)
    tok = _tokenizer(d, kw, 16)
    tok.update_stats(wav.cuda(), lengths.cuda())
    own = tok.mel_transform(wav.cuda()).cpu()
    lo_own, hi_own = O.calibrate_minmax(own, n_valid)
    assert torch.equal(tok.quantizer.lo.cpu(), lo_own) and torch.equal(tok.quantizer.hi.cpu(), hi_own)
    lo_ref, hi_ref = O.calibrate_minmax(mel_ref, n_valid)
    tok.quantizer.set_stats(lo_ref, hi_ref)
    codes, code_lengths = tok.encode(wav.cuda(), lengths.cuda())
    assert torch.equal(code_lengths.cpu().long(), n_valid)
    _check_codes(codes, mel_ref, lo_ref, hi_ref, 16, n_valid)
    # tensor-level calibration on an existing log-mel gives the same stats
    q = d.DMelQuantizer(80, 16).cuda()
    q.update_stats(own.cuda(), n_valid.cuda())
    assert torch.equal(q.lo.cpu(), lo_own) and torch.equal(q.hi.cpu(), hi_own)


def test_calibration_composes_over_batches(d):
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    wav = synth.batch(range(400, 408), 16000, 16000, "speech").cuda()
    whole, parts = _tokenizer(d, kw, 16), _tokenizer(d, kw, 16)
    whole.calibrate([wav])
    parts.calibrate([wav[:3], wav[3:5], (wav[5:], None)])
    assert torch.equal(whole.quantizer.lo, parts.quantizer.lo) and torch.equal(whole.quantizer.hi, parts.quantizer.hi)


def test_encode_host_matches_device_path(d):
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg2_24k_128"]
    wav = synth.batch(range(500, 537), 48000, 24000, "speech")  # 37 rows -> several chunks? no: rows are small
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav.cuda()])
    want, _ = tok.encode(wav.cuda())
    got = tok.encode_host(wav.pin_memory())
    assert not got.is_cuda and torch.equal(got, want.cpu())
    lengths = torch.randint(300, 48000, (37,), dtype=torch.int32)
    want, _ = tok.encode(wav.cuda(), lengths.cuda())
    assert torch.equal(tok.encode_host(wav, lengths), want.cpu())


def test_encode_host_follows_a_change_of_statistics(d):
    """the library keeps the statistics it last uploaded and skips the upload when a call brings the same values:
    new values (another calibration, another tokenizer on the same plan geometry) must reach the device"""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    wav = synth.batch(range(600, 606), 32000, 16000, "speech")
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav.cuda()])
    first = tok.encode_host(wav)
    assert torch.equal(first, tok.encode(wav.cuda())[0].cpu())
    assert torch.equal(tok.encode_host(wav), first)                      # same statistics: the cached upload
    q = tok.quantizer
    q.set_stats(q.lo - 1.5, q.hi + 0.75)
    second = tok.encode_host(wav)
    assert torch.equal(second, tok.encode(wav.cuda())[0].cpu()) and not torch.equal(second, first)
    q.set_stats(q.lo + 1.5, q.hi - 0.75)                                  # and back
    assert torch.equal(tok.encode_host(wav), first)


def test_quantizer_edge_cases(d):
    q = d.DMelQuantizer(3, 16).cuda()
    q.set_stats(torch.tensor([-11.5, 0.0, 2.0]), torch.tensor([1.0, 0.0, 3.0]))  # channel 1 is degenerate
    x = torch.tensor([[[-20.0, -11.5, 1.0, 5.0], [0.0, 0.0, 0.0, 7.0], [2.0, 2.5, 3.0, 2.999]]], device="cuda")
    codes = q.encode(x)
    want = O.dmel_encode(x.cpu(), q.lo.cpu(), q.hi.cpu(), 16)
    assert torch.equal(codes.cpu(), want)
    assert codes[0, 0].tolist() == [0, 0, 15, 15] and codes[0, 1].tolist() == [0, 0, 0, 0]
    assert torch.equal(q.decode(codes).cpu(), O.dmel_decode(want, q.lo.cpu(), q.hi.cpu(), 16))
    # out-of-range codes clamp to the last bin instead of reading past the table
    wild = torch.full((1, 3, 5), 200, dtype=torch.uint8, device="cuda")
    assert torch.equal(q.decode(wild), q.decode(torch.full_like(wild, 15)))
    # empty batch
    assert q.encode(torch.zeros(0, 3, 4, device="cuda")).shape == (0, 3, 4)


@pytest.mark.parametrize("shape", [(1, 80, 1), (3, 80, 5), (2, 128, 937), (5, 100, 61), (1, 160, 5167), (1, 1, 5), (7, 3, 1),
                                   (2, 64, 1024), (3, 1, 4097)])
def test_streaming_kernels_odd_shapes(d, shape):
    """Flat-indexed quantise / dequantise / min-max cross row boundaries inside a 4-element group."""
    b, m, t = shape
    g = torch.Generator().manual_seed(b * 1000 + t)
    mel = (torch.rand(shape, generator=g) * 12 - 11.5)
    q = d.DMelQuantizer(m, 16).cuda()
    q.update_stats(mel.cuda())
    lo, hi = O.calibrate_minmax(mel)
    assert torch.equal(q.lo.cpu(), lo) and torch.equal(q.hi.cpu(), hi)
    codes = q.encode(mel.cuda())
    assert torch.equal(codes.cpu(), O.dmel_encode(mel, lo, hi, 16))
    assert torch.equal(q.decode(codes).cpu(), O.dmel_decode(codes.cpu(), lo, hi, 16))
    # misaligned views fall back to the scalar path and must agree
    buf = torch.zeros(mel.numel() + 1, device="cuda")
    buf[1:] = mel.flatten().cuda()
    assert torch.equal(q.encode(buf[1:].view(shape)), codes)


# ---------------------------------------------------------------------------
# BASELINE.json full-size configuration: size-independent properties
# ---------------------------------------------------------------------------
def test_full_size_config2_properties(d):
    """configs[1]: 24 kHz, 128 mel, 16 bins, batch 64 x 10 s.  The oracle is too slow to run
    at this size inside the suite, so check properties: a sampled sub-batch against the
    oracle, determinism, decode(encode) within half a bin, tiling independence."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg2_24k_128"]
    b, n = 64, 240000
    wav = synth.device_batch(range(b), n, 24000, "cuda")
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav])
    codes, _, mel = tok.encode(wav, return_mel=True)
    assert codes.shape == (b, 128, 937)
    codes2, _ = tok.encode(wav)
    assert torch.equal(codes, codes2)  # deterministic
    # batch tiling independence: rows encoded alone give the same bytes
    alone, _ = tok.encode(wav[17:18])
    assert torch.equal(alone[0], codes[17])
    # decode(encode(x)) within half a bin of the kernel's own log-mel
    back = tok.decode(codes)
    half = ((tok.quantizer.hi - tok.quantizer.lo) / 16 / 2)[None, :, None]
    assert torch.all((back - mel).abs() <= half * (1 + 1e-4) + 1e-6)
    # both ends of every channel's range are used
    assert codes.amin(dim=(0, 2)).eq(0).all() and codes.amax(dim=(0, 2)).eq(15).all()
    # sampled rows against the CPU oracle
    rows = [0, 31, 63]
    ref = O.log_mel(wav[rows].cpu(), oracle_config(kw))
    ok, ratio = logmel_close(mel[rows].cpu(), ref, REL_TOL)
    assert ok, f"{ratio:.2f}x tolerance"
    _check_codes(codes[rows], ref, tok.quantizer.lo.cpu(), tok.quantizer.hi.cpu(), 16)


# ---------------------------------------------------------------------------
# streaming (BASELINE configs[3]): chunked output == offline output, bit for bit
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("chunking", ["80ms", "ragged", "one_shot"])
@pytest.mark.parametrize("n_streams", [1, 3])
def test_streaming_equals_offline(d, chunking, n_streams):
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    n = 16000 * 3 + 123
    wav = synth.batch(range(700, 700 + n_streams), n, 16000, "speech").cuda()
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav])
    want, _ = tok.encode(wav)
    enc = d.DMelStreamEncoder(tok, n_streams=n_streams, capacity_samples=8192)
    if chunking == "80ms":
        sizes = [1280] * (n // 1280) + ([n % 1280] if n % 1280 else [])
    elif chunking == "ragged":
        g = torch.Generator().manual_seed(3)
        sizes, left = [], n
        while left > 0:
            c = min(left, int(torch.randint(1, 3000, (1,), generator=g)))
            sizes.append(c)
            left -= c
    else:
        enc = d.DMelStreamEncoder(tok, n_streams=n_streams, capacity_samples=n + 8)
        sizes = [n]
    parts, at, counts = [], 0, []
    for c in sizes:
        out = enc.push(wav[:, 0, at:at + c])
        at += c
        parts.append(out)
        counts.append(out.shape[2])
    parts.append(enc.flush())
    got = torch.cat(parts, dim=2)
    assert got.shape == want.shape and torch.equal(got, want)
    if chunking == "80ms":
        assert counts[0] == 3 and all(k == 5 for k in counts[1:-1])  # 3 frames, then 5 per chunk
    # the encoder is reusable after flush()
    again = torch.cat([enc.push(wav[:, 0, :4000]), enc.flush()], dim=2)
    assert torch.equal(again, tok.encode(wav[:, :, :4000])[0])


def test_streaming_host_chunks_and_errors(d):
    """Chunks may come straight from host memory; misuse raises like the offline path."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    n = 16000 + 77
    wav = synth.batch(range(710, 712), n, 16000, "speech")
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav.cuda()])
    want, _ = tok.encode(wav.cuda())
    enc = d.DMelStreamEncoder(tok, n_streams=2, capacity_samples=4096)
    parts = [enc.push(wav[:, 0, at:at + 1280].contiguous()) for at in range(0, n, 1280)]  # CPU tensors
    parts.append(enc.flush())
    assert torch.equal(torch.cat(parts, dim=2), want)
    with pytest.raises(ValueError):
        enc.flush()  # nothing received since the reset: shorter than the reflect pad
    with pytest.raises(ValueError):
        enc.push(torch.zeros(3, 10))  # wrong number of streams
    with pytest.raises(ValueError):
        enc.push(torch.zeros(2, 1 << 20))  # does not fit the history buffer


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_masked_logmel_matches_call_site(d, dtype):
    """LogMelSpectrogram.masked == transform -> .to(dtype) -> * sequence_mask of the reference's
    VQGAN.encode_unquantized (models/codec_lit_modules.py:486-507), in one launch."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg2_24k_128"]
    n = 24000 * 2 + 300
    wav = synth.batch(range(720, 725), n, 24000, "speech").cuda()
    lengths = torch.tensor([n, n - 1, 24000, 700, 0], dtype=torch.int64, device="cuda")
    mt = d.LogMelSpectrogram(**kw).cuda()
    full = mt(wav)                                   # (B, M, T) float32, every frame
    mels, mel_lengths = mt.masked(wav, lengths, dtype)
    assert mels.dtype == dtype and mels.shape == full.shape
    assert torch.equal(mel_lengths, lengths // kw["hop_length"])
    mask = (torch.arange(full.shape[2], device="cuda")[None, :] < mel_lengths[:, None])[:, None, :]
    want = full.to(dtype) * mask.to(dtype)
    assert torch.equal(mels, want)                   # same bits (0 == -0 under torch.equal)
    # the group view the caller takes is a view, not a copy
    g = 8
    assert mels.view(mels.shape[0] * g, mels.shape[1] // g, mels.shape[2]).data_ptr() == mels.data_ptr()
    # no lengths: plain cast
    assert torch.equal(mt.spectrogram.plan_for(wav.device).logmel_masked(wav, None, dtype), full.to(dtype))


def test_dynamic_tile_schedule_matches_static(d, monkeypatch):
    """Tiles handed out by the global counter give the same bytes as the static walk, launch after
    launch (the counter pair of a launch must be back at zero when its slot comes round again)."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg2_24k_128"]
    wav = synth.device_batch(range(40), 24000 * 4 + 77, 24000, "cuda")
    lengths = torch.randint(1000, wav.shape[-1], (40,), generator=torch.Generator().manual_seed(5)).to("cuda", torch.int32)
    tok = _tokenizer(d, kw, 16)
    tok.calibrate([wav])
    monkeypatch.setenv("DMEL_STATIC_TILES", "1")
    want, _ = tok.encode(wav, lengths)
    want_mel = tok.mel_transform(wav)
    monkeypatch.delenv("DMEL_STATIC_TILES")
    for _ in range(70):  # more launches than counter slots
        got, _ = tok.encode(wav, lengths)
        assert torch.equal(got, want)
    assert torch.equal(tok.mel_transform(wav), want_mel)


@pytest.mark.parametrize("name,n_bins", [("cfg2_24k_128", 16), ("cfg1_16k_80", 16), ("cfg5_44k_160", 32)])
def test_pcm16_input_matches_float_path(d, name, n_bins):
    """int16 PCM in (device and host entry points) gives the same bytes as the float32 path on
    sample / 32768, including ragged lengths, odd row lengths and a strided batch."""
    kw = GOLDEN_GEOMETRY[name]
    n = kw["sample_rate"] * 2 + 1237
    g = torch.Generator().manual_seed(17)
    pcm = (torch.randn(6, n, generator=g) * 6000).clamp(-32768, 32767).to(torch.int16)
    pcm[1, 5000:9000] = 0                       # digital silence: the clamp floor
    pcm[2] = torch.randint(-32768, 32768, (n,), generator=g).to(torch.int16)  # full scale
    lengths = torch.tensor([n, n - 1, n // 2, 3000, n, 1], dtype=torch.int32)
    tok = _tokenizer(d, kw, n_bins)
    as_float = pcm.float() / 32768.0
    tok.calibrate([as_float.cuda()])
    want, want_len = tok.encode(as_float.cuda(), lengths.cuda())
    got, got_len = tok.encode_pcm16(pcm.cuda(), lengths.cuda())
    assert torch.equal(got, want) and torch.equal(got_len, want_len)
    assert torch.equal(tok.encode_pcm16(pcm.cuda()[:, None, :])[0], tok.encode(as_float.cuda())[0])   # (B, 1, L), no lengths
    wide = torch.zeros(6, n + 40, dtype=torch.int16, device="cuda")
    wide[:, :n] = pcm.cuda()
    assert torch.equal(tok.encode_pcm16(wide[:, :n], lengths.cuda())[0], want)    # row stride > row length
    host = tok.encode_host(pcm.pin_memory(), lengths)                              # int16 host buffers
    assert torch.equal(host, want.cpu())
    with pytest.raises(ValueError):
        tok.encode_pcm16(as_float.cuda())


def test_quality_statistic_from_fused_sums(d):
    """LogMelSpectrogram.with_quality == the reference training step's quality on the same mel
    (models/codec_lit_modules.py:171-174), from sums the mel launch accumulates itself."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY["cfg2_24k_128"]
    wav = synth.batch(range(730, 738), 24000 * 3 + 11, 24000, "speech").cuda()
    wav[5] *= 1e-3   # a quiet utterance: fewer channels above the -8 threshold
    mt = d.LogMelSpectrogram(**kw).cuda()
    mels, quality = mt.with_quality(wav)
    assert torch.equal(mels, mt(wav))
    mean = mels.mean(-1)
    want = ((mean > -8).sum(-1) - 90) / 10
    near = ((mean + 8).abs() < 1e-4).any(-1)          # a channel mean sitting on the threshold may flip
    assert quality.shape == (8, 1)
    assert torch.equal(quality[~near, 0], want[~near].to(quality.dtype))
    assert len(set(quality[:, 0].tolist())) > 1          # the statistic discriminates in this batch


@pytest.mark.parametrize("name,n_bins", [("cfg2_24k_128", 16), ("cfg5_44k_160", 32)])
def test_fused_forward_equals_encode_then_decode(d, name, n_bins):
    """DMelTokenizer.encode_decode (one launch) == encode followed by the table-lookup decode, bit for bit,
    with and without lengths."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY[name]
    n = kw["sample_rate"] * 2 + 333
    wav = synth.batch(range(740, 745), n, kw["sample_rate"], "speech").cuda()
    lengths = torch.tensor([n, n // 2, 5000, n - 1, 0], device="cuda")
    tok = _tokenizer(d, kw, n_bins)
    tok.calibrate([wav])
    codes, code_lengths = tok.encode(wav, lengths)
    out = tok.encode_decode(wav, lengths)
    assert torch.equal(out.codes, codes)
    assert torch.equal(out.z, tok.decode(codes, code_lengths))
    codes2, _ = tok.encode(wav)
    out2 = tok.encode_decode(wav)
    assert torch.equal(out2.codes, codes2) and torch.equal(out2.z, tok.decode(codes2))


def _log_mel_float64(wav, cfg):
    """The reference's chain (utils/spectrogram.py:41-81) evaluated in float64 on the same float32
    window and filterbank: the yardstick for the rounding error of both float32 implementations."""
    w = wav.squeeze(1).double() if wav.ndim == 3 else wav.double()
    padded = torch.nn.functional.pad(w[:, None, :], (cfg.pad, cfg.pad), mode="reflect")[:, 0, :]
    frames = padded.unfold(-1, cfg.n_fft, cfg.hop_length)
    spec = torch.fft.rfft(frames * O.stft_window(cfg.win_length, cfg.n_fft).double(), dim=-1)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9).transpose(1, 2)
    bank = torch.from_numpy(O.slaney_filterbank(cfg.sample_rate, cfg.n_fft, cfg.n_mels, cfg.f_min, cfg.f_max)).double()
    return torch.log(torch.clamp(torch.matmul(bank, mag), min=1e-5))


@pytest.mark.parametrize("name", ["cfg2_24k_128", "cfg5_44k_160"])
def test_rounding_error_against_float64_is_no_worse_than_the_reference(d, name):
    """Against a float64 evaluation, the kernel's log-mel error stays within twice the error of the
    reference's own float32 path (torch.stft + matmul on the CPU) on speech-like and Gaussian audio."""
    from dmel_codec_b200 import synth
    kw = GOLDEN_GEOMETRY[name]
    cfg = oracle_config(kw)
    n = kw["sample_rate"] * 2 + 57
    wav = torch.cat([synth.batch(range(750, 753), n, kw["sample_rate"], "speech"),
                     synth.batch(range(753, 755), n, kw["sample_rate"], "noise")])
    truth = _log_mel_float64(wav, cfg)
    ref = O.log_mel(wav, cfg).double()
    got = _transform(d, kw)(wav.cuda()).cpu().double()
    scale = torch.clamp(truth.abs(), min=1.0)
    err_ref = ((ref - truth).abs() / scale).max().item()
    err_got = ((got - truth).abs() / scale).max().item()
    print(f"{name}: max error vs float64  kernel {err_got:.2e}  reference float32 path {err_ref:.2e}")
    assert err_got <= max(2.0 * err_ref, 5e-6), (err_got, err_ref)
    assert err_got <= 1e-4


def test_single_transform_calibrate_encode_equals_two_pass_job(d):
    """distributed.calibrate_encode_sharded (log-mel stored by the calibration pass, stand-alone quantiser as
    pass 2) gives the statistics and the codes of calibrate_sharded + encode_sharded, bit for bit."""
    from dmel_codec_b200 import distributed as D, synth
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    n = 16000 * 2 + 101
    pool = synth.batch(range(760, 772), n, 16000, "speech").cuda()
    lengths = torch.randint(2000, n, (12,), generator=torch.Generator().manual_seed(9)).cuda()
    for with_lengths in (False, True):
        load = (lambda ids: (pool[list(ids)], lengths[list(ids)])) if with_lengths else (lambda ids: pool[list(ids)])
        a, b = _tokenizer(d, kw, 16), _tokenizer(d, kw, 16)
        D.calibrate_sharded(a, 12, load, 5)
        want = list(D.encode_sharded(a, 12, load, 5))
        got = list(D.calibrate_encode_sharded(b, 12, load, 5))
        assert torch.equal(a.quantizer.lo, b.quantizer.lo) and torch.equal(a.quantizer.hi, b.quantizer.hi)
        assert len(want) == len(got) == 3
        for (ids_w, codes_w, len_w), (ids_g, codes_g, len_g) in zip(want, got):
            assert list(ids_w) == list(ids_g) and torch.equal(codes_w, codes_g)
            assert (len_w is None and len_g is None) or torch.equal(len_w, len_g)
