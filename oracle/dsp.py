"""Oracle DSP primitives (torch CPU, fp32) -- test infrastructure, see oracle/__init__.py.

Each function restates one torchaudio / torch call site of the reference.
``TA:`` = torchaudio (site-packages), ``TORCH:`` = torch (site-packages).
"""
from __future__ import annotations

import math

import torch


def hann_periodic(n_fft: int) -> torch.Tensor:
    """``torch.hann_window(n_fft)`` as passed via ``window_fn`` (app3.py:135-139).

    Periodic Hann: w[n] = 0.5 - 0.5 cos(2 pi n / n_fft).
    """
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)


def reflect_pad(x: torch.Tensor, p: int) -> torch.Tensor:
    """``pad_mode="reflect"`` of torch.stft (TORCH:functional.py:675-690).

    xp = [x[p], ..., x[1], x[0..L-1], x[L-2], ..., x[L-1-p]].
    """
    if p == 0:
        return x
    left = x[..., 1 : p + 1].flip(-1)
    right = x[..., -p - 1 : -1].flip(-1)
    return torch.cat([left, x, right], dim=-1)


def stft(x: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """Complex spectrogram, TA:functional/functional.py:54-145 with
    ``power=None, center=True, pad_mode="reflect", normalized=False, onesided=True``
    (call sites app3.py:191, server.py:207).

    x: [..., L] -> [..., F, T] complex64, F = n_fft//2+1, T = 1 + L//hop.
    """
    w = hann_periodic(n_fft)
    xp = reflect_pad(x, n_fft // 2)
    frames = xp.unfold(-1, n_fft, hop) * w  # [..., T, n_fft]
    spec = torch.fft.rfft(frames, n=n_fft, dim=-1)  # [..., T, F]
    return spec.transpose(-1, -2).contiguous()


def istft(spec: torch.Tensor, n_fft: int, hop: int, length: int | None = None) -> torch.Tensor:
    """``torch.istft`` as reached from TA:functional/functional.py:205 (InverseSpectrogram,
    server.py:216) and from the Griffin-Lim loop (TA:functional/functional.py:316-348).

    spec: [..., F, T] complex -> [..., hop*(T-1)] (``length=None``).
    frame_t = irfft(spec[:, t]) * w ; overlap-add ; divide by sum of w^2 ; trim n_fft//2.
    """
    w = hann_periodic(n_fft)
    T = spec.shape[-1]
    frames = torch.fft.irfft(spec.transpose(-1, -2), n=n_fft, dim=-1) * w  # [..., T, n_fft]
    total = n_fft + hop * (T - 1)
    lead = frames.shape[:-2]
    acc = torch.zeros(*lead, total, dtype=frames.dtype)
    env = torch.zeros(total, dtype=frames.dtype)
    w2 = w * w
    for t in range(T):
        acc[..., t * hop : t * hop + n_fft] += frames[..., t, :]
        env[t * hop : t * hop + n_fft] += w2
    p = n_fft // 2
    end = total - p if length is None else p + length
    out = acc[..., p:end] / env[p:end]
    return out


def mel_fbanks(n_freqs: int, n_mels: int, sample_rate: int) -> torch.Tensor:
    """HTK triangular filterbank, ``norm=None``, ``f_min=0``, ``f_max=sample_rate//2``:
    TA:functional/functional.py:518-588 + :492-515, as built by ``MelScale``
    (app3.py:140-143).  Returns fb [n_freqs, n_mels] float32.
    """
    f_max = float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)  # [n_freqs, n_mels+2]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def mel_scale(mag: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """``MelScale.forward`` TA:transforms/_transforms.py:407-419: [..., F, T] -> [..., M, T]."""
    return torch.matmul(mag.transpose(-1, -2), fb).transpose(-1, -2)


def log_mel(x: torch.Tensor, n_fft: int, hop: int, fb: torch.Tensor) -> torch.Tensor:
    """``M0T(T0(x).abs()).log1p()`` (app3.py:191-193, server.py:207-210) -> [..., M, T]."""
    return mel_scale(stft(x, n_fft, hop).abs(), fb).log1p()


def inverse_mel(mel: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """``InverseMelScale.forward`` TA:transforms/_transforms.py:491-512 (driver "gels"):
    relu(min-norm solution of fb^T X = mel).  mel [..., M, T] -> [..., F, T].
    """
    shape = mel.shape
    m = mel.reshape(-1, shape[-2], shape[-1])
    sol = torch.linalg.lstsq(fb.transpose(-1, -2)[None], m, driver="gels").solution
    return torch.relu(sol).reshape(shape[:-2] + sol.shape[-2:])


def inverse_mel_pinv(fb: torch.Tensor) -> torch.Tensor:
    """P = pinv(fb^T) [F, M] in fp64 -> fp32.  relu(P @ mel) equals inverse_mel to ~2e-7
    (SURVEY.md K5); this is the fixed matrix the CUDA path multiplies by."""
    return torch.linalg.pinv(fb.double().transpose(0, 1)).float().contiguous()


def griffinlim(
    mag: torch.Tensor,
    n_fft: int,
    hop: int,
    n_iter: int = 32,
    momentum: float = 0.99,
    init_angles: torch.Tensor | None = None,
    rand_init: bool = True,
) -> torch.Tensor:
    """Fast Griffin-Lim, ``power=1`` : TA:functional/functional.py:255-353 as configured at
    app3.py:149-153 (defaults n_iter=32, momentum=0.99, rand_init=True, length=None).

    ``init_angles`` (complex, same shape as ``mag``) replaces the ``torch.rand`` draw at
    TA:functional/functional.py:310 so that two implementations can be compared.
    mag: [..., F, T] -> [..., hop*(T-1)].
    """
    if not 0 <= momentum < 1:
        raise ValueError(f"momentum must be in range [0, 1). Found: {momentum}")
    m = momentum / (1 + momentum)
    shape = mag.shape
    mag = mag.reshape(-1, shape[-2], shape[-1])
    if init_angles is not None:
        angles = init_angles.reshape(mag.shape).to(torch.complex64)
    elif rand_init:
        angles = torch.rand(mag.shape, dtype=torch.complex64)
    else:
        angles = torch.ones(mag.shape, dtype=torch.complex64)
    prev = None
    for _ in range(n_iter):
        wave = istft(mag * angles, n_fft, hop)
        rebuilt = stft(wave, n_fft, hop)
        angles = rebuilt if (prev is None or m == 0) else rebuilt - m * prev
        angles = angles / (angles.abs() + 1e-16)
        prev = rebuilt
    wave = istft(mag * angles, n_fft, hop)
    return wave.reshape(shape[:-2] + wave.shape[-1:])
