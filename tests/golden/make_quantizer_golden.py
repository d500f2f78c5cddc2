#!/usr/bin/env python
"""Freeze the dMel quantiser spec as golden vectors: tests/golden/quantizer_golden.npz.

    python tests/golden/make_quantizer_golden.py

The reference ships no bin quantiser, so this spec is the repo's own; what this file pins is that nobody
changes it silently.  The arithmetic below is written out in plain numpy float32 — deliberately NOT an import of
oracle/ or of the product — and both the oracle and the CUDA path are tested against the stored outputs, bit
for bit (tests/test_oracle.py, tests/test_gpu_parity.py).

Inputs are the reference-generated log-mels of tests/golden/reference_logmel.npz (made by make_golden.py from
the reference's own utils/spectrogram.py) plus one hand-built tensor with the edge cases: a degenerate channel
(hi == lo), values equal to hi and to lo, values exactly on interior edges lo + i*step, values one float32 ulp on
either side of an edge, values outside [lo, hi], and the clamp floor.
"""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32
CASES = {"cfg1_16k_80": 16, "cfg2_24k_128": 16, "cfg5_44k_160": 32, "yaml_24k_100": 16, "short_window": 256}


def calibrate(mel):
    """per-channel min / max over batch and time of a (B, M, T) float32 tensor"""
    return mel.min(axis=(0, 2)).astype(F), mel.max(axis=(0, 2)).astype(F)


def spec(mel, lo, hi, k):
    """Appendix B, every operation a separately rounded float32 operation:
         w = hi - lo;  s = K / w  (0 if w <= 0);  code = clamp(floor((x - lo) * s), 0, K - 1)
         step = w / K; table[c][j] = lo + (j + 0.5) * step;  x_hat = table[c][code]"""
    w = (hi - lo).astype(F)
    s = np.where(w > 0, F(k) / np.where(w > 0, w, F(1)), F(0)).astype(F)
    d = (mel - lo[None, :, None]).astype(F)
    pos = (d * s[None, :, None]).astype(F)
    codes = np.clip(np.floor(pos), 0, k - 1).astype(np.uint8)
    step = (w / F(k)).astype(F)
    centre = (np.arange(k, dtype=F) + F(0.5)).astype(F)
    prod = (centre[None, :] * step[:, None]).astype(F)
    table = (lo[:, None] + prod).astype(F)
    decoded = table[np.arange(mel.shape[1])[None, :, None], codes.astype(np.int64)]
    return s, step, table, codes, decoded.astype(F)


def edge_case_tensor(k=16):
    """(1, 6, 4k+8): channel 0 degenerate, channels 1-5 with [lo, hi] of different magnitudes; values on, just
    below and just above every interior edge, at lo, at hi, outside the range."""
    ranges = [(F(-3.25), F(-3.25)), (F(-11.512925), F(1.5)), (F(-8.0), F(0.0)), (F(-0.1), F(0.1)),
              (F(-11.512925), F(-11.0)), (F(0.5), F(7.75))]
    rows = []
    for lo, hi in ranges:
        step = F((hi - lo) / F(k))
        vals = [lo, hi, F(lo - F(1.0)), F(hi + F(1.0)), F(-11.512925148010254), F(lo + step / F(2)), F(hi - step / F(2)), F(0.0)]
        for i in range(1, k + 1):
            e = F(lo + F(i) * step)
            vals += [e, np.nextafter(e, F(-np.inf), dtype=F), np.nextafter(e, F(np.inf), dtype=F), F(e - step / F(3))]
        rows.append(np.array(vals, dtype=F))
    x = np.stack(rows)[None]
    lo = np.array([r[0] for r in ranges], dtype=F)
    hi = np.array([r[1] for r in ranges], dtype=F)
    return x, lo, hi


def main():
    g = np.load(os.path.join(HERE, "reference_logmel.npz"))
    out = {}
    for name, k in CASES.items():
        mel = g[name + "/logmel"].astype(F)
        lo, hi = calibrate(mel)
        s, step, table, codes, decoded = spec(mel, lo, hi, k)
        out.update({f"{name}/n_bins": np.int32(k), f"{name}/lo": lo, f"{name}/hi": hi, f"{name}/scale": s,
                    f"{name}/step": step, f"{name}/table": table, f"{name}/codes": codes, f"{name}/decoded": decoded})
        hist = np.bincount(codes.ravel(), minlength=k)
        print(f"{name}: K={k} codes {codes.shape}, used bins {int((hist > 0).sum())}/{k}")
    x, lo, hi = edge_case_tensor(16)
    s, step, table, codes, decoded = spec(x, lo, hi, 16)
    out.update({"edges/n_bins": np.int32(16), "edges/mel": x, "edges/lo": lo, "edges/hi": hi, "edges/scale": s,
                "edges/step": step, "edges/table": table, "edges/codes": codes, "edges/decoded": decoded})
    print("edges: degenerate channel codes", np.unique(codes[0, 0]), "| codes at value == hi:", codes[0, 1:, 1])
    path = os.path.join(HERE, "quantizer_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
