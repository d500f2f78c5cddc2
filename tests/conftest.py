import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libdmel_b200.so."""
    from dmel_codec_b200 import _native, build
    build.build()
    return _native.load()


@pytest.fixture(scope="session")
def golden():
    """Reference log-mel fixtures written by tests/golden/make_golden.py."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_logmel.npz"))


# geometry of every golden case: ctor kwargs of LogMelSpectrogram (mirrors make_golden.CASES)
GOLDEN_GEOMETRY = {
    "cfg1_16k_80": dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
    "cfg2_24k_128": dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=128, f_min=0, f_max=12000),
    "cfg5_44k_160": dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=160),
    "yaml_24k_100": dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100, f_min=0, f_max=12000),
    "edge_len_hop_minus1": dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
    "edge_len_hop_plus1": dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
    "edge_len_min": dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
    "edge_len_1280": dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
    "short_window": dict(sample_rate=22050, n_fft=1024, win_length=800, hop_length=200, n_mels=64),
    "silence": dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
}


def oracle_config(kw):
    from oracle import dmel_oracle as O
    return O.MelConfig(sample_rate=kw["sample_rate"], n_fft=kw["n_fft"], win_length=kw["win_length"],
                       hop_length=kw["hop_length"], n_mels=kw["n_mels"], f_min=kw.get("f_min", 0.0),
                       f_max=kw.get("f_max"), center=kw.get("center", False))


def logmel_close(a, b, rel=1e-4):
    """BASELINE.md section 6 gate: |a-b| <= rel * max(1, |b|). Returns (ok, worst ratio)."""
    import torch
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    ratio = ((a - b).abs() / (rel * torch.clamp(b.abs(), min=1.0))).max().item()
    return ratio <= 1.0, ratio
