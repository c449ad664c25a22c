"""Deterministic, reference-independent parameter fill used by the golden
fixtures: the same function fills the reference module's state_dict (in
make_golden.py, container only) and the native module's / oracle's state_dict
(in the tests), so no multi-MB weight file has to be committed.

Each tensor gets its own generator seeded by crc32(key) ^ seed, so the result
does not depend on key order.
"""
import zlib
import math
import torch


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _mel_init(n, sr=16000, min_low=50, min_band=50):
    to_mel = lambda hz: 2595.0 * math.log10(1 + hz / 700.0)
    lo, hi = to_mel(30.0), to_mel(sr / 2 - (min_low + min_band))
    mel = torch.linspace(lo, hi, n + 1, dtype=torch.float64)
    hz = 700.0 * (10 ** (mel / 2595.0) - 1)
    return hz[:-1].float().view(-1, 1), (hz[1:] - hz[:-1]).float().view(-1, 1)


def fill_state_dict(sd, seed=0):
    """Returns a NEW dict with the same keys/shapes/dtypes, deterministic values."""
    out = {}
    for key in sorted(sd.keys()):
        ref = sd[key]
        g = _gen(key, seed)
        shape = tuple(ref.shape)
        leaf = key.split(".")[-1]
        if leaf == "num_batches_tracked":
            v = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "low_hz_":
            v = _mel_init(shape[0])[0] + 3.0 * torch.randn(shape, generator=g)
        elif leaf == "band_hz_":
            v = _mel_init(shape[0])[1] + 3.0 * torch.randn(shape, generator=g)
        elif leaf == "running_mean":
            v = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            v = 0.5 + torch.rand(shape, generator=g)
        elif ".norm." in key and leaf == "weight":
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".act." in key and leaf == "weight":
            v = 0.2 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            v = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) >= 2:
            if "deconv" in key:                     # (Cin, Cout, k): fan-in Cin*k/stride-ish
                fan = shape[0] * shape[2] / 4.0
            else:
                fan = 1
                for d in shape[1:]:
                    fan *= d
            v = torch.randn(shape, generator=g) / math.sqrt(max(fan, 1.0))
        else:
            v = 0.1 * torch.randn(shape, generator=g)
        out[key] = v.to(ref.dtype)
    return out


def seeded_randn(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return scale * torch.randn(shape, generator=g)


def sample_view(t, stride=997):
    """Deterministic strided subsample used to pin very large gradients."""
    return t.reshape(-1)[::stride].clone()
