#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It loads ``/root/reference/dmel_codec/utils/spectrogram.py`` by path, unmodified.
That file imports ``librosa.filters.mel`` (librosa 0.10.2.post1, not installed
and not installable offline), so one stand-in module is injected into
``sys.modules`` first; the stand-in is the oracle's restatement of that
function.  Everything else — reflect pad, ``torch.stft``, magnitude, matmul,
log-clamp — is the reference's own arithmetic.

The fixtures hold the seeded input waveforms and the reference's log-mel output
for small cases of each BASELINE geometry plus the edge lengths the survey
lists, so they travel to the GPU box where /root/reference does not exist.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import dmel_oracle  # noqa: E402
from dmel_codec_b200 import synth  # noqa: E402

REF_FILE = "/root/reference/dmel_codec/utils/spectrogram.py"


def load_reference():
    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
        return dmel_oracle.slaney_filterbank(sr, n_fft, n_mels, fmin, fmax)

    librosa = types.ModuleType("librosa")
    filters = types.ModuleType("librosa.filters")
    filters.mel = mel
    librosa.filters = filters
    sys.modules.setdefault("librosa", librosa)
    sys.modules.setdefault("librosa.filters", filters)
    spec = importlib.util.spec_from_file_location("_ref_spectrogram", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# name -> (ctor kwargs for LogMelSpectrogram, batch, n_samples, kind, first utterance id)
CASES = {
    # BASELINE configs[0] geometry (16 kHz / 1024 / 256 / 80), 2 s so the silent span covers whole frames
    "cfg1_16k_80": (dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80), 2, 32000, "speech", 0),
    # configs[1] geometry (24 kHz / 1024 / 256 / 128, f_max 12000 as the reference's yaml passes it)
    "cfg2_24k_128": (dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=128, f_min=0, f_max=12000), 2, 12000, "speech", 10),
    # configs[4] geometry (44.1 kHz / 2048 / 512 / 160): reference defaults + 160 mel
    "cfg5_44k_160": (dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=160), 2, 22050, "speech", 20),
    # the reference's shipped 24 kHz codec config uses 100 mel channels (config/codec/dMel_used.yaml:20-24)
    "yaml_24k_100": (dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100, f_min=0, f_max=12000), 2, 6000, "noise", 30),
    # edge lengths: L = k*hop +- 1, L just above the reflect minimum, streaming chunk size
    "edge_len_hop_minus1": (dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80), 2, 256 * 9 - 1, "noise", 40),
    "edge_len_hop_plus1": (dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80), 2, 256 * 9 + 1, "noise", 42),
    "edge_len_min": (dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80), 1, 385, "noise", 44),
    "edge_len_1280": (dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80), 1, 1280, "speech", 45),
    # a window shorter than n_fft and a hop that does not divide n_fft
    "short_window": (dict(sample_rate=22050, n_fft=1024, win_length=800, hop_length=200, n_mels=64), 2, 5000, "speech", 50),
    # digital silence: every value must be the clamp floor
    "silence": (dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80), 1, 4096, "zeros", 0),
}


def make_input(batch, n, sr, kind, first_id):
    if kind == "zeros":
        return torch.zeros(batch, 1, n)
    return synth.batch(range(first_id, first_id + batch), n, sr, kind)


def main():
    ref = load_reference()
    torch.manual_seed(0)
    out = {}
    for name, (kw, b, n, kind, uid) in CASES.items():
        wav = make_input(b, n, kw["sample_rate"], kind, uid)
        tr = ref.LogMelSpectrogram(**kw)
        with torch.no_grad():
            mel = tr(wav)
        out[name + "/wav"] = wav.numpy()
        out[name + "/logmel"] = mel.numpy().astype(np.float32)
        print(f"{name}: wav {tuple(wav.shape)} -> logmel {tuple(mel.shape)}  "
              f"min {mel.min():.4f} max {mel.max():.4f}")
    path = os.path.join(HERE, "reference_logmel.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
