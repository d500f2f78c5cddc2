#!/usr/bin/env python
"""Generate tests/golden/activation_golden.npz by running the REFERENCE's own anti-aliased activation modules.

    python tests/golden/make_activation_golden.py        (build container only: needs /root/reference)

Imports, unmodified, dmel_codec/models/modules/bigvgan/alias_free_activation/torch/{act,resample,filter}.py and
dmel_codec/models/modules/bigvgan/activations.py from /root/reference, builds Activation1d(SnakeBeta / Snake) and stores
seeded inputs, parameters and outputs for small cases (incl. sequences shorter than the filter, which are all padding)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from dmel_codec.models.modules.bigvgan import activations  # noqa: E402
from dmel_codec.models.modules.bigvgan.alias_free_activation.torch.act import Activation1d  # noqa: E402

CASES = {  # name -> (B, C, T, activation class name, parameter scale)
    "snakebeta_small": (2, 6, 50, "SnakeBeta", 0.5),
    "snakebeta_long": (1, 4, 1337, "SnakeBeta", 1.0),
    "snake_shared_param": (2, 3, 64, "Snake", 0.7),
    "shorter_than_filter": (1, 2, 5, "SnakeBeta", 0.3),
    "single_sample": (1, 2, 1, "SnakeBeta", 0.3),
    "large_arguments": (1, 3, 200, "SnakeBeta", 2.0),  # exp(alpha) up to ~50: sin arguments far outside [-pi, pi]
}


def main():
    out = {}
    for seed, (name, (b, c, t, cls, scale)) in enumerate(CASES.items()):
        g = torch.Generator().manual_seed(4200 + seed)
        act = getattr(activations, cls)(c, alpha_logscale=True)
        with torch.no_grad():
            act.alpha.copy_(torch.randn(c, generator=g) * scale * 2)
            if cls == "SnakeBeta":
                act.beta.copy_(torch.randn(c, generator=g) * scale)
        module = Activation1d(activation=act)
        x = torch.randn(b, c, t, generator=g) * 1.5
        with torch.no_grad():
            y = module(x)
        out[name + "/x"] = x.numpy()
        out[name + "/log_alpha"] = act.alpha.detach().numpy()
        out[name + "/log_beta"] = (act.beta if cls == "SnakeBeta" else act.alpha).detach().numpy()
        out[name + "/up_taps"] = module.upsample.filter.reshape(-1).numpy()
        out[name + "/down_taps"] = module.downsample.lowpass.filter.reshape(-1).numpy()
        out[name + "/y"] = y.numpy()
        print(f"{name}: x {tuple(x.shape)} -> y {tuple(y.shape)}  |y| max {y.abs().max():.3f}")
    path = os.path.join(HERE, "activation_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
