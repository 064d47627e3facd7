"""Oracle compositions of the hot path -- test infrastructure, see oracle/__init__.py.

* ``denoise_batch``      : app3.py:181-217 applied to whole clips.
* ``denoise_noisy_phase``: server.py:207-216 (noisy-phase iSTFT, no Griffin-Lim).
* ``StreamingOracle``    : app3.py:123-250 ``DenoisingAudioProcessor`` hop loop.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import dsp


def peak_normalise(x: torch.Tensor):
    """app3.py:181-186 per clip: divide by max|x| when it exceeds 1e-6, else leave (peak := 1)."""
    peak = x.abs().amax(dim=-1, keepdim=True)
    ok = peak > 1e-6
    peak = torch.where(ok, peak, torch.ones_like(peak))
    return x / peak, peak


@torch.no_grad()
def denoise_batch(
    noisy: torch.Tensor,
    model,
    n_fft: int = 1024,
    hop: int = 512,
    n_mels: int = 64,
    sample_rate: int = 16000,
    n_iter: int = 32,
    momentum: float = 0.99,
    init_angles: torch.Tensor | None = None,
    hx: torch.Tensor | None = None,
    normalise: bool = True,
) -> dict:
    """noisy [B, L] -> dict(wave [B, hop*(T-1)], logmel, pred, mel_mag, lin_mag, hx, peak)."""
    fb = dsp.mel_fbanks(n_fft // 2 + 1, n_mels, sample_rate)
    if normalise:
        x, peak = peak_normalise(noisy)
    else:
        x, peak = noisy, torch.ones(noisy.shape[:-1] + (1,))
    logmel = dsp.log_mel(x, n_fft, hop, fb)  # [B, M, T]      app3.py:191-193
    feats = logmel.transpose(-1, -2)  # [B, T, M]      app3.py:195
    pred, hx = model(feats, hx)  # app3.py:201
    rec = F.leaky_relu(feats - pred, negative_slope=0.2)  # app3.py:203-204
    mel_mag = torch.expm1(rec.transpose(-1, -2)).clamp(min=0)  # app3.py:206-208
    lin_mag = dsp.inverse_mel(mel_mag, fb).clamp(min=0)  # app3.py:210-211
    wave = dsp.griffinlim(lin_mag, n_fft, hop, n_iter, momentum, init_angles)  # app3.py:213
    wave = wave * peak  # app3.py:217
    return dict(wave=wave, logmel=logmel, pred=pred, mel_mag=mel_mag, lin_mag=lin_mag, hx=hx, peak=peak)


@torch.no_grad()
def denoise_noisy_phase(
    x: torch.Tensor,
    model,
    n_fft: int = 1024,
    hop: int = 512,
    n_mels: int = 64,
    sample_rate: int = 48000,
    hx: torch.Tensor | None = None,
    out_scale: float = 3.0,
    hx_decay: float = 0.9,
) -> dict:
    """server.py:207-216: x [B, L] -> wave [B, hop*(T-1)] using the noisy phase."""
    fb = dsp.mel_fbanks(n_fft // 2 + 1, n_mels, sample_rate)
    spec = dsp.stft(x, n_fft, hop)
    phase = spec.angle()
    logmel = dsp.mel_scale(spec.abs(), fb).log1p()
    pred, hx = model(logmel.transpose(-1, -2), hx)
    out = F.leaky_relu(pred.transpose(-1, -2), negative_slope=0) * out_scale  # server.py:213
    hx = hx * hx_decay  # server.py:214
    lin = dsp.inverse_mel((logmel - out).exp() - 1, fb)  # server.py:215
    wave = dsp.istft(torch.polar(lin, phase), n_fft, hop)  # server.py:216
    return dict(wave=wave, logmel=logmel, pred=pred, lin_mag=lin, hx=hx)


class StreamingOracle:
    """Hop loop of ``DenoisingAudioProcessor.recv`` (app3.py:167-226) on float32 chunks.

    Quirks Q2-Q4 of SURVEY.md Appendix C are kept: the chunk is Hann-windowed before the
    (centre-padded, windowed again) STFT, so each hop produces 3 frames and 3 GRU steps,
    and the Griffin-Lim output is overlap-added without a synthesis window, one hop late.
    ``angles_fn(hop_index, shape)`` supplies the Griffin-Lim initial angles.
    """

    def __init__(self, model, n_fft=1536, hop=768, n_mels=64, sample_rate=48000, n_iter=32, angles_fn=None):
        self.model = model
        self.n_fft, self.hop, self.n_mels, self.sr, self.n_iter = n_fft, hop, n_mels, sample_rate, n_iter
        self.fb = dsp.mel_fbanks(n_fft // 2 + 1, n_mels, sample_rate)
        self.win = dsp.hann_periodic(n_fft).numpy()
        self.inbuf = np.zeros(0, dtype=np.float32)
        self.ola = np.zeros(n_fft, dtype=np.float32)
        self.hx = None
        self.hops = 0
        self.angles_fn = angles_fn

    @torch.no_grad()
    def push(self, chunk: np.ndarray) -> np.ndarray:
        self.inbuf = np.concatenate([self.inbuf, chunk.astype(np.float32)])
        produced = np.zeros(0, dtype=np.float32)
        n_fft, hop = self.n_fft, self.hop
        while len(self.inbuf) >= n_fft:
            cur = self.inbuf[:n_fft]
            peak = np.max(np.abs(cur))
            if peak > 1e-6:
                cur = cur / peak
            else:
                peak = 1.0
            x = torch.from_numpy(cur * self.win).float().unsqueeze(0)
            logmel = dsp.log_mel(x, n_fft, hop, self.fb)
            feats = logmel.transpose(-1, -2)
            pred, self.hx = self.model(feats, self.hx)
            rec = F.leaky_relu(feats - pred, negative_slope=0.2)
            mel_mag = torch.expm1(rec.transpose(-1, -2)).clamp(min=0)
            lin = dsp.inverse_mel(mel_mag, self.fb).clamp(min=0)
            init = None if self.angles_fn is None else self.angles_fn(self.hops, lin.shape)
            y = dsp.griffinlim(lin, n_fft, hop, self.n_iter, 0.99, init).squeeze(0).numpy() * peak
            produced = np.concatenate([produced, self.ola[:hop].copy()])
            self.ola[:-hop] = self.ola[hop:]
            self.ola[-hop:] = 0.0
            self.ola[:n_fft] += y
            self.inbuf = self.inbuf[hop:]
            self.hops += 1
        return produced
