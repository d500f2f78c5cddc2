"""dmel_codec_b200 — B200-native dMel tokenization hot path.

Public surface (drop-in for the reference's ``dmel_codec.utils.spectrogram``
plus the dMel quantiser the reference's README describes but does not ship):

    LinearSpectrogram, LogMelSpectrogram      waveform -> log-mel
    DMelQuantizer, DMelResult                 log-mel <-> uint8 codes, calibration
    DMelTokenizer                             waveform -> codes (fused kernel)
    DMelStreamEncoder                         the same, chunk by chunk (bit-identical to offline)
    FSQIndexer                                next to the path: the weight-free core of the reference's learned
                                              quantiser (FSQ codes / indices / id_shift)
    AntiAliasSnake                            next to the path: BigVGAN's anti-aliased Snake activation, one kernel

Everything computes in hand-written sm_100a CUDA reached through the C ABI in
``include/dmel_b200.h``; importing the package is cheap, the first call loads
``libdmel_b200.so`` and raises if it was not built.
"""
from .spectrogram import LinearSpectrogram, LogMelSpectrogram
from .quantizer import DMelQuantizer, DMelResult, DMelTokenizer
from .streaming import DMelStreamEncoder
from .fsq import FSQIndexer
from .activation import AntiAliasSnake

__all__ = ["LinearSpectrogram", "LogMelSpectrogram", "DMelQuantizer", "DMelResult", "DMelTokenizer",
           "DMelStreamEncoder", "FSQIndexer", "AntiAliasSnake"]
