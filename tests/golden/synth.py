"""Deterministic, platform-independent pseudo-random tensors for fixtures that store no weights.

Values come from an integer hash of (seed, element index) evaluated with int64 tensor arithmetic (wrap-around keeps
the low 32 bits exact), so a generator script run in the build container and a test run anywhere produce identical
bf16-representable tensors.  Used by tests/golden/make_golden_qwen.py, make_golden_wan.py and the tests that read
their fixtures.
"""
import zlib

import torch

_M32 = 0xFFFFFFFF


def _hash32(x):
    x = (x ^ (x >> 16)) * 0x45D9F3B & _M32
    x = (x ^ (x >> 16)) * 0x45D9F3B & _M32
    return (x ^ (x >> 16)) & _M32


def synth_uniform(n, seed):
    """n values in [-0.5, 0.5), float64."""
    idx = torch.arange(n, dtype=torch.int64)
    x = _hash32((idx * 0x9E3779B1 + (int(seed) & _M32) * 0x85EBCA6B + 0x27D4EB2F) & _M32)
    return x.to(torch.float64) / 4294967296.0 - 0.5


def synth_tensor(shape, seed, std=1.0, mean=0.0):
    """fp32 tensor of the given shape, zero-mean uniform with standard deviation `std` (+ mean), rounded to bf16 grid."""
    n = 1
    for s in shape:
        n *= int(s)
    v = synth_uniform(n, seed) * (std * 12 ** 0.5) + mean
    return v.to(torch.float32).to(torch.bfloat16).to(torch.float32).reshape(shape)


def name_seed(name):
    return zlib.crc32(name.encode())


def fill_parameters(module, w_std=0.05, b_std=0.05):
    """Fills every parameter by name: matrices ~ std 0.05, norm scales ~ 1 + 0.1 u, other vectors ~ std 0.05."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.ndim > 1:
                v = synth_tensor(tuple(p.shape), name_seed(name), w_std)
            elif 'norm' in name and name.endswith('weight'):
                v = synth_tensor(tuple(p.shape), name_seed(name), 0.1, 1.0)
            else:
                v = synth_tensor(tuple(p.shape), name_seed(name), b_std)
            p.copy_(v.to(p.dtype))
    return module


def fingerprint(t, seed):
    """(sum, abs-sum, dot with a fixed probe, max-abs) of a tensor — enough to pin a gradient without storing it."""
    if t is None:
        return None
    t = t.detach().double().flatten()
    probe = synth_uniform(t.numel(), seed)
    return {'sum': float(t.sum()), 'abs': float(t.abs().sum()), 'dot': float((t * probe).sum()), 'max': float(t.abs().max()),
            'numel': t.numel(), 'seed': int(seed)}


def fingerprint_close(fp, t, rtol, what=''):
    """Compares a tensor against a stored fingerprint; tolerances are relative to the abs-sum (the natural scale)."""
    if fp is None:                      # the reference produced no gradient (parameter not on the path to the loss)
        assert t is None or float(t.abs().max()) == 0.0, what
        return
    assert t is not None, what
    got = fingerprint(t, fp['seed'])
    assert got['numel'] == fp['numel'], (what, got['numel'], fp['numel'])
    scale = max(fp['abs'], 1e-12)
    for k in ('sum', 'dot'):
        assert abs(got[k] - fp[k]) <= rtol * scale, (what, k, got[k], fp[k], scale)
    assert abs(got['abs'] - fp['abs']) <= rtol * scale, (what, 'abs', got['abs'], fp['abs'])
