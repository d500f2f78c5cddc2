"""The reference arm: ishine/dmel_codec's own, unmodified code, for bench.py and the comparison tests.

``install()`` (called by ``__graft_entry__.build()`` in the build container, where ``/root/reference`` exists)
pip-installs the reference into ``baseline/_ref`` — git-ignored, but it travels to the GPU box with the
snapshot:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference>

``--no-deps`` because ``librosa``, ``lightning``, ``hydra``, ``vector_quantize_pytorch``, ``lhotse`` … are neither
installed nor in the wheelhouse; the hot path's file, ``dmel_codec/utils/spectrogram.py``, needs only torch,
torchaudio and ``librosa.filters.mel`` (reference utils/spectrogram.py:1-4).  ``load()`` injects a stand-in for
that one function and imports ``dmel_codec.utils.spectrogram`` from ``baseline/_ref`` — the reference's own class
behind the reference's own import path, the string its Hydra configs bind (config/codec/dMel_used.yaml:88).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
REF_FILE = os.path.join(REF_DIR, "dmel_codec", "utils", "spectrogram.py")


def available() -> bool:
    return os.path.exists(REF_FILE)


def install(force: bool = False) -> str:
    """Install the reference package into baseline/_ref (no-op when it is already there or the sources are absent)."""
    if available() and not force and os.path.isdir(os.path.join(REF_DIR, "dmel_codec", "models", "modules", "bigvgan")):
        return "present"
    if not os.path.isdir(REF_SRC):
        return "reference sources absent (GPU box): using what travelled with the snapshot"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "reference")  # /root/reference is read-only and setuptools writes into the tree
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", REF_DIR, "--upgrade", src]
        res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or not available():
        return f"pip install failed ({res.returncode}): {res.stderr.strip()[-300:]}"
    # The reference's setup.py finds packages by __init__.py, and models/modules/bigvgan has none: pip leaves the vocoder
    # (and with it the anti-aliased activation, torch path and CUDA sources) out.  Complete the install from the sources.
    extra = os.path.join("dmel_codec", "models", "modules", "bigvgan")
    dst = os.path.join(REF_DIR, extra)
    if not os.path.isdir(dst):
        shutil.copytree(os.path.join(REF_SRC, extra), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return "installed"


def load_activation():
    """The reference's anti-aliased activation, torch path: (Activation1d class, activations module), imported
    unmodified from baseline/_ref (namespace packages: the vocoder directory has no __init__.py)."""
    path = os.path.join(REF_DIR, "dmel_codec", "models", "modules", "bigvgan", "alias_free_activation", "torch", "act.py")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: run __graft_entry__.build() where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    act = importlib.import_module("dmel_codec.models.modules.bigvgan.alias_free_activation.torch.act")
    activations = importlib.import_module("dmel_codec.models.modules.bigvgan.activations")
    return act.Activation1d, activations


def load(mel_fn):
    """Import the reference's ``dmel_codec.utils.spectrogram`` from baseline/_ref.

    ``mel_fn(sr, n_fft, n_mels, fmin, fmax) -> (n_mels, n_fft//2+1) float32 ndarray`` stands in for
    ``librosa.filters.mel`` (librosa 0.10.2.post1 is not installable offline)."""
    if not available():
        raise FileNotFoundError(f"{REF_FILE} not found: run __graft_entry__.build() where /root/reference exists")

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
        return mel_fn(sr, n_fft, n_mels, fmin, fmax)

    librosa = types.ModuleType("librosa")
    filters = types.ModuleType("librosa.filters")
    filters.mel = mel
    librosa.filters = filters
    sys.modules["librosa"] = librosa
    sys.modules["librosa.filters"] = filters
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return importlib.import_module("dmel_codec.utils.spectrogram")
