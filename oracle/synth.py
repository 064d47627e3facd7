"""Seeded synthetic clips -- test infrastructure, see oracle/__init__.py.

clean = tone + chirp + speech-like harmonic stack; noisy = clean + white noise at a
per-clip SNR drawn from U[0, 20] dB, then peak-normalised to 1.0 (app3.py:181-186).
Everything is generated on the host from ``torch.Generator().manual_seed(1234 + clip)``.
"""
from __future__ import annotations

import math

import torch


def make_clip(index: int, length: int, sr: int = 16000) -> tuple[torch.Tensor, torch.Tensor]:
    """Returns (noisy[L], clean[L]) float32, both scaled by the noisy clip's peak."""
    g = torch.Generator().manual_seed(1234 + index)
    u = torch.rand(6, generator=g, dtype=torch.float64)
    t = torch.arange(length, dtype=torch.float64) / sr
    dur = max(length / sr, 1e-9)
    f0 = 100.0 + 1900.0 * u[0].item()
    f1 = 100.0 + 400.0 * u[1].item()
    f2 = 1000.0 + 5000.0 * u[2].item()
    snr_db = 20.0 * u[3].item()
    tone = 0.5 * torch.sin(2 * math.pi * f0 * t)
    chirp = 0.5 * torch.sin(2 * math.pi * (f1 + (f2 - f1) * t / (2 * dur)) * t)
    pitch = 120.0 + 30.0 * torch.sin(2 * math.pi * 3.0 * t)
    voiced = torch.zeros_like(t)
    for k in range(1, 11):
        voiced = voiced + torch.sin(2 * math.pi * k * pitch * t) / k
    am = 0.5 + 0.5 * torch.sin(2 * math.pi * 4.0 * t)
    syll = (torch.sin(2 * math.pi * 3.0 * t + 2 * math.pi * u[4].item()) > 0).double()
    clean = tone + chirp + 0.3 * voiced * am * syll
    noise = torch.randn(length, generator=g, dtype=torch.float64)
    p_clean = clean.pow(2).mean().clamp_min(1e-12)
    p_noise = noise.pow(2).mean().clamp_min(1e-12)
    noise = noise * torch.sqrt(p_clean / (p_noise * 10.0 ** (snr_db / 10.0)))
    noisy = clean + noise
    peak = noisy.abs().max().clamp_min(1e-6)
    return (noisy / peak).float(), (clean / peak).float()


def make_batch(n: int, length: int, sr: int = 16000, start: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    pairs = [make_clip(start + i, length, sr) for i in range(n)]
    return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])


def make_batch_fast(n: int, length: int, sr: int = 16000, seed: int = 1234) -> torch.Tensor:
    """Cheap bench-sized noisy batch: a handful of real clips tiled with per-clip gain/shift.

    Content does not affect the timing of any kernel on the path (no data-dependent
    control flow), so bench.py uses this to avoid minutes of host-side synthesis.
    """
    base, _ = make_batch(min(n, 8), length, sr)
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(n, length, dtype=torch.float32)
    for i in range(n):
        src = base[i % base.shape[0]]
        shift = int(torch.randint(0, length, (1,), generator=g))
        noise = 0.05 * torch.randn(length, generator=g)
        clip = torch.roll(src, shift) + noise
        out[i] = clip / clip.abs().max().clamp_min(1e-6)
    return out


def gl_init_angles(shape, seed: int = 7) -> torch.Tensor:
    """The draw of TA:functional/functional.py:310 under a fixed seed: real & imag ~ U[0,1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, dtype=torch.complex64, generator=g)
