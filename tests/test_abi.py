"""CPU tests of the boundary: the shared library loads, exports exactly what
include/dmel_b200.h declares, and refuses to work without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from dmel_codec_b200 import _native


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "dmel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmel_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(_native.SIGNATURES)


def test_library_exports_every_declared_symbol(native_lib):
    for name in _declared_functions():
        assert hasattr(native_lib, name), name
    assert native_lib.dmel_abi_version() == _native.ABI_VERSION


def test_extension_is_built_for_sm100a():
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(native_lib):
    import numpy as np
    handle = ctypes.c_void_p()
    basis = np.zeros((80, 513), np.float32)
    window = np.ones(1024, np.float32)
    rc = native_lib.dmel_plan_create(1024, 256, 80, 0, basis.ctypes.data_as(ctypes.c_void_p),
                                     window.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle))
    assert rc == _native.ERR_NO_DEVICE and not handle.value
    assert b"no CPU path" in native_lib.dmel_last_error()
    with pytest.raises(_native.DmelNativeError):
        _native.check(rc)


def test_argument_errors_without_touching_a_device(native_lib):
    import numpy as np
    handle = ctypes.c_void_p()
    basis = np.zeros((80, 257), np.float32)
    window = np.ones(512, np.float32)
    rc = native_lib.dmel_plan_create(400, 100, 80, 0, basis.ctypes.data_as(ctypes.c_void_p),  # not a power of two
                                     window.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle))
    assert rc == _native.ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _native.check(rc)
    rc = native_lib.dmel_plan_create(1024, 0, 80, 0, basis.ctypes.data_as(ctypes.c_void_p),
                                     window.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle))
    assert rc == _native.ERR_INVALID
    with pytest.raises(ValueError):
        _native.check(rc)
    assert native_lib.dmel_quantize_u8(None, 1, 80, 10, None, None, 16, None, None) == _native.ERR_INVALID
    assert native_lib.dmel_run(None, None, None) == _native.ERR_INVALID
    bad_levels = (ctypes.c_int * 3)(7, 1, 5)
    assert native_lib.dmel_fsq_encode(None, 1, 10, 10, bad_levels, 3, None, None, None, 0, None) == _native.ERR_INVALID


def test_modules_reject_cpu_tensors():
    import dmel_codec_b200 as d
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.LogMelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80)(torch.zeros(1, 4096))
    q = d.DMelQuantizer(80, 16)
    q.set_stats(torch.zeros(80), torch.ones(80))
    with pytest.raises(RuntimeError, match="GPU only"):
        q.encode(torch.zeros(1, 80, 4))
    with pytest.raises(RuntimeError, match="no calibration"):
        d.DMelQuantizer(80, 16).encode(torch.zeros(1, 80, 4))


def test_drop_in_constructor_surface():
    """Same keywords, defaults and attributes as reference utils/spectrogram.py:9-20, :85-117."""
    import inspect
    import dmel_codec_b200 as d
    sig = inspect.signature(d.LogMelSpectrogram.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [
        ("sample_rate", 44100), ("n_fft", 2048), ("win_length", 2048), ("hop_length", 512), ("n_mels", 128),
        ("center", False), ("f_min", 0.0), ("f_max", None)]
    sig = inspect.signature(d.LinearSpectrogram.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [
        ("n_fft", 2048), ("win_length", 2048), ("hop_length", 512), ("center", False), ("num_mels", 128),
        ("f_min", 0), ("f_max", None), ("sample_rate", 44100), ("mode", "reflect")]
    m = d.LogMelSpectrogram(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100, f_max=12000)
    assert (m.hop_length, m.sample_rate, m.n_fft, m.win_length, m.n_mels, m.center) == (256, 24000, 1024, 1024, 100, False)
    assert m.f_max == 12000 and d.LogMelSpectrogram().f_max == 22050.0 and d.LogMelSpectrogram().spectrogram.f_max is None
    assert list(m.state_dict()) == []  # like the reference: no buffers, no parameters
    assert sorted(d.DMelTokenizer(n_mels=80).state_dict()) == ["quantizer.hi", "quantizer.lo"]
