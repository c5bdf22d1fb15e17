"""Deterministic synthetic weights for golden fixtures that store OUTPUTS only (test infrastructure).

``fill(sd, seed)`` overwrites every floating-point tensor of a state_dict with values that depend only on (seed, key,
shape): the fixture generator (which has the reference) and the test (which has not) build identical weights, so a
mid-size configuration can be pinned without committing megabytes of parameters."""
import zlib

import torch


def fill(sd, seed=0):
    out = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            out[k] = v.clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) & 0x7FFFFFFF)
        t = torch.randn(v.shape, generator=g, dtype=torch.float32)
        if k.endswith("running_var"):
            t = t.abs() * 0.5 + 0.5
        elif k.endswith("running_mean") or k.endswith(".bias"):
            t = t * 0.1
        elif v.dim() == 1:                       # BatchNorm scale
            t = 1.0 + 0.1 * t
        else:                                    # convolution weights: keep activations O(1) through ~60 layers
            fan_in = v[0].numel() if v.dim() > 1 else 1
            t = t * (0.9 / fan_in ** 0.5)
        out[k] = t
    return out
