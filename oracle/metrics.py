"""Parity metrics -- test infrastructure, see oracle/__init__.py."""
from __future__ import annotations

import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b|| / ||b|| over the whole tensor (complex ok)."""
    a = a.detach().cpu()
    b = b.detach().cpu()
    if a.is_complex() or b.is_complex():
        a = torch.view_as_real(a.to(torch.complex128))
        b = torch.view_as_real(b.to(torch.complex128))
    a = a.double()
    b = b.double()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def si_sdr(est: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """Scale-invariant SDR in dB along the last axis; one value per leading index."""
    est = est.detach().cpu().double()
    ref = ref.detach().cpu().double()
    alpha = (est * ref).sum(-1, keepdim=True) / ref.pow(2).sum(-1, keepdim=True).clamp_min(1e-30)
    target = alpha * ref
    noise = est - target
    return 10.0 * torch.log10(target.pow(2).sum(-1).clamp_min(1e-30) / noise.pow(2).sum(-1).clamp_min(1e-30))


def spectral_convergence(mag_est: torch.Tensor, mag_ref: torch.Tensor) -> float:
    """|| |STFT(y)| - mag || / || mag ||."""
    return rel_l2(mag_est, mag_ref)
