"""Drop-in replacements for the five ``torchaudio.transforms`` the reference constructs
(app3.py:135-153, server.py:173-176), backed by the sm_100a kernels of libb200denoise.so.

Same constructor keywords and call conventions as torchaudio (leading batch dims are packed,
``.to(device)`` works because they are ``nn.Module``s); options the reference never uses raise
``NotImplementedError`` instead of silently falling back to a library path.  CUDA tensors only.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from . import _cabi
from ._runtime import Workspace, draw_seed, get_plan, is_hann, ptr, require_cuda_c64, require_cuda_f32, stream_ptr

__all__ = ["Spectrogram", "MelScale", "InverseMelScale", "GriffinLim", "InverseSpectrogram", "Resample", "pcm16_to_float", "float_to_pcm16"]


def _geometry(n_fft, win_length, hop_length, window_fn, wkwargs, who):
    win_length = win_length if win_length is not None else n_fft
    hop_length = hop_length if hop_length is not None else win_length // 2
    if win_length != n_fft:
        raise NotImplementedError(f"{who}: win_length must equal n_fft (the reference's only setting), got {win_length} vs {n_fft}")
    if not is_hann(window_fn, win_length, wkwargs):
        raise NotImplementedError(f"{who}: only the periodic Hann window (torch.hann_window) is implemented")
    return n_fft, hop_length


def _pack(x: torch.Tensor, keep: int):
    """[..., d1..dkeep] -> [B, d1..dkeep]"""
    lead = x.shape[:-keep]
    return x.reshape((-1,) + tuple(x.shape[-keep:])), lead


class Spectrogram(nn.Module):
    """``torchaudio.transforms.Spectrogram(power=None, ...)`` (app3.py:135-139, server.py:173).

    forward(waveform [..., L]) -> complex64 [..., n_fft//2+1, 1 + L//hop].
    """

    def __init__(self, n_fft: int = 400, win_length: Optional[int] = None, hop_length: Optional[int] = None, pad: int = 0,
                 window_fn: Callable[..., torch.Tensor] = torch.hann_window, power: Optional[float] = 2.0,
                 normalized: bool = False, wkwargs: Optional[dict] = None, center: bool = True, pad_mode: str = "reflect",
                 onesided: bool = True) -> None:
        super().__init__()
        if power is not None:
            raise NotImplementedError("Spectrogram: only power=None (complex output) is implemented, as used by the reference")
        if pad != 0 or normalized or not center or pad_mode != "reflect" or not onesided:
            raise NotImplementedError("Spectrogram: only pad=0, normalized=False, center=True, pad_mode='reflect', onesided=True")
        self.n_fft, self.hop_length = _geometry(n_fft, win_length, hop_length, window_fn, wkwargs, "Spectrogram")
        self.win_length = self.n_fft
        self.power = None

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        x = require_cuda_f32(waveform, "waveform")
        x, lead = _pack(x, 1)
        B, L = x.shape
        plan = get_plan(self.n_fft, self.hop_length, 0, 0, x.device)
        T = plan.num_frames(L)
        spec = torch.empty((B, plan.n_freqs, T), dtype=torch.complex64, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().b2d_stft(plan.handle, x.data_ptr(), B, L, spec.data_ptr(), stream_ptr(x.device)))
        return spec.reshape(lead + spec.shape[-2:])


class MelScale(nn.Module):
    """``torchaudio.transforms.MelScale`` (app3.py:140-143): [..., n_stft, T] -> [..., n_mels, T]."""

    def __init__(self, n_mels: int = 128, sample_rate: int = 16000, f_min: float = 0.0, f_max: Optional[float] = None,
                 n_stft: int = 201, norm: Optional[str] = None, mel_scale: str = "htk") -> None:
        super().__init__()
        if f_min != 0.0 or (f_max is not None and float(f_max) != float(sample_rate // 2)) or norm is not None or mel_scale != "htk":
            raise NotImplementedError("MelScale: only f_min=0, f_max=sample_rate//2, norm=None, mel_scale='htk' are implemented")
        self.n_mels, self.sample_rate, self.n_stft = n_mels, sample_rate, n_stft
        self.f_min, self.f_max = f_min, float(sample_rate // 2)

    def _plan(self, device):
        n_fft = 2 * (self.n_stft - 1)
        return get_plan(n_fft, n_fft // 2, self.n_mels, self.sample_rate, device)

    @property
    def fb(self) -> torch.Tensor:
        from ._runtime import melscale_fbanks_htk

        return melscale_fbanks_htk(self.n_stft, self.n_mels, self.sample_rate)

    def forward(self, specgram: torch.Tensor) -> torch.Tensor:
        s = require_cuda_f32(specgram, "specgram")
        s, lead = _pack(s, 2)
        B, F, T = s.shape
        if F != self.n_stft:
            raise ValueError(f"Expected an input with {self.n_stft} frequency bins. Found: {F}")
        plan = self._plan(s.device)
        mel = torch.empty((B, self.n_mels, T), dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            _cabi.check(_cabi.lib().b2d_mel_scale(plan.handle, s.data_ptr(), B, T, mel.data_ptr(), stream_ptr(s.device)))
        return mel.reshape(lead + mel.shape[-2:])


class InverseMelScale(nn.Module):
    """``torchaudio.transforms.InverseMelScale`` (app3.py:145-148): relu(lstsq(fb^T, mel)) computed as
    relu(pinv(fb^T) @ mel) (identical minimum-norm solution).  [..., n_mels, T] -> [..., n_stft, T]."""

    def __init__(self, n_stft: int, n_mels: int = 128, sample_rate: int = 16000, f_min: float = 0.0,
                 f_max: Optional[float] = None, norm: Optional[str] = None, mel_scale: str = "htk", driver: str = "gels") -> None:
        super().__init__()
        if f_min != 0.0 or (f_max is not None and float(f_max) != float(sample_rate // 2)) or norm is not None or mel_scale != "htk":
            raise NotImplementedError("InverseMelScale: only f_min=0, f_max=sample_rate//2, norm=None, mel_scale='htk' are implemented")
        if driver not in ["gels", "gelsy", "gelsd", "gelss"]:
            raise ValueError(f'driver must be one of ["gels", "gelsy", "gelsd", "gelss"]. Found {driver}.')
        self.n_mels, self.sample_rate, self.n_stft, self.driver = n_mels, sample_rate, n_stft, driver
        self.f_min, self.f_max = f_min, float(sample_rate // 2)

    def _plan(self, device):
        n_fft = 2 * (self.n_stft - 1)
        plan = get_plan(n_fft, n_fft // 2, self.n_mels, self.sample_rate, device)
        if plan.rank < self.n_mels:
            raise ValueError(
                f"InverseMelScale: the mel filterbank for n_stft={self.n_stft}, n_mels={self.n_mels}, sample_rate={self.sample_rate} "
                f"has rank {plan.rank} < n_mels; the reference's lstsq(driver='gels') is undefined there (SURVEY.md section 7.3)"
            )
        return plan

    def forward(self, melspec: torch.Tensor) -> torch.Tensor:
        m = require_cuda_f32(melspec, "melspec")
        m, lead = _pack(m, 2)
        B, n_mels, T = m.shape
        if self.n_mels != n_mels:
            raise ValueError("Expected an input with {} mel bins. Found: {}".format(self.n_mels, n_mels))
        plan = self._plan(m.device)
        lin = torch.empty((B, self.n_stft, T), dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            _cabi.check(_cabi.lib().b2d_inverse_mel(plan.handle, m.data_ptr(), B, T, lin.data_ptr(), stream_ptr(m.device)))
        return lin.reshape(lead + lin.shape[-2:])


class GriffinLim(nn.Module):
    """``torchaudio.transforms.GriffinLim`` (app3.py:149-153).  forward(specgram [..., F, T]) -> [..., hop*(T-1)].

    ``forward(specgram, init_angles=...)`` injects the initial complex ``angles`` tensor that torchaudio draws
    with ``torch.rand`` (TA:functional/functional.py:310) so two implementations can be compared.
    """

    def __init__(self, n_fft: int = 400, n_iter: int = 32, win_length: Optional[int] = None, hop_length: Optional[int] = None,
                 window_fn: Callable[..., torch.Tensor] = torch.hann_window, power: float = 2.0, wkwargs: Optional[dict] = None,
                 momentum: float = 0.99, length: Optional[int] = None, rand_init: bool = True) -> None:
        super().__init__()
        if not (0 <= momentum < 1):
            raise ValueError("momentum must be in the range [0, 1). Found: {}".format(momentum))
        if length is not None:
            raise NotImplementedError("GriffinLim: only length=None is implemented (the reference's setting)")
        self.n_fft, self.hop_length = _geometry(n_fft, win_length, hop_length, window_fn, wkwargs, "GriffinLim")
        if self.hop_length * 2 != self.n_fft:
            raise NotImplementedError("GriffinLim: only hop_length == n_fft // 2 is implemented (the reference's setting)")
        self.win_length = self.n_fft
        self.n_iter, self.power, self.momentum, self.length, self.rand_init = n_iter, power, momentum, length, rand_init
        self._ws = Workspace()

    def forward(self, specgram: torch.Tensor, init_angles: Optional[torch.Tensor] = None) -> torch.Tensor:
        s = require_cuda_f32(specgram, "specgram")
        s, lead = _pack(s, 2)
        if self.power != 1.0:
            s = s.pow(1.0 / self.power)
        B, F, T = s.shape
        if F != self.n_fft // 2 + 1:
            raise ValueError(f"Expected {self.n_fft // 2 + 1} frequency bins. Found: {F}")
        seed = 0
        if init_angles is None and self.rand_init:
            # rand_init=True (TA:functional/functional.py:310): U[0,1) real/imag parts drawn inside the kernel from a
            # counter-based generator seeded from torch's CPU generator (so torch.manual_seed makes runs repeatable)
            seed = draw_seed()
        if init_angles is not None:
            init_angles = require_cuda_c64(init_angles, "init_angles").reshape(s.shape)
        plan = get_plan(self.n_fft, self.hop_length, 0, 0, s.device)
        lib = _cabi.lib()
        nbytes = lib.b2d_griffinlim_workspace_bytes(plan.handle, B, T)
        ws = self._ws.get(nbytes, s.device)
        wave = torch.empty((B, plan.out_length(T)), dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            _cabi.check(lib.b2d_griffinlim(plan.handle, s.data_ptr(), ptr(init_angles), seed, B, T, self.n_iter, float(self.momentum),
                                           None, wave.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(s.device)))
        return wave.reshape(lead + wave.shape[-1:])


class InverseSpectrogram(nn.Module):
    """``torchaudio.transforms.InverseSpectrogram`` (server.py:174,216): complex [..., F, T] -> [..., hop*(T-1)]."""

    def __init__(self, n_fft: int = 400, win_length: Optional[int] = None, hop_length: Optional[int] = None, pad: int = 0,
                 window_fn: Callable[..., torch.Tensor] = torch.hann_window, normalized: bool = False, wkwargs: Optional[dict] = None,
                 center: bool = True, pad_mode: str = "reflect", onesided: bool = True) -> None:
        super().__init__()
        if pad != 0 or normalized or not center or pad_mode != "reflect" or not onesided:
            raise NotImplementedError("InverseSpectrogram: only pad=0, normalized=False, center=True, onesided=True")
        self.n_fft, self.hop_length = _geometry(n_fft, win_length, hop_length, window_fn, wkwargs, "InverseSpectrogram")
        if self.hop_length * 2 != self.n_fft:
            raise NotImplementedError("InverseSpectrogram: only hop_length == n_fft // 2 is implemented (the reference's setting)")
        self.win_length = self.n_fft

    def forward(self, spectrogram: torch.Tensor, length: Optional[int] = None, magnitude: Optional[torch.Tensor] = None) -> torch.Tensor:
        if length is not None:
            raise NotImplementedError("InverseSpectrogram: only length=None is implemented")
        s = require_cuda_c64(spectrogram, "spectrogram")
        s, lead = _pack(s, 2)
        B, F, T = s.shape
        if F != self.n_fft // 2 + 1:
            raise ValueError(f"Expected {self.n_fft // 2 + 1} frequency bins. Found: {F}")
        mag = None
        if magnitude is not None:  # istft(polar(magnitude, angle(spectrogram))) -- server.py:216
            mag = require_cuda_f32(magnitude, "magnitude").reshape(s.shape)
        plan = get_plan(self.n_fft, self.hop_length, 0, 0, s.device)
        wave = torch.empty((B, plan.out_length(T)), dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            _cabi.check(_cabi.lib().b2d_istft(plan.handle, s.data_ptr(), ptr(mag), B, T, wave.data_ptr(), stream_ptr(s.device)))
        return wave.reshape(lead + wave.shape[-1:])


def _sinc_hann_table(orig: int, new: int, lowpass_filter_width: int, rolloff: float):
    """Polyphase table of TA:functional/functional.py:1340-1402 (sinc_interp_hann): float64 arithmetic, float32 result
    [new, 2*width + orig]."""
    import math

    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = torch.arange(0, -new, -1)[:, None] / new + idx  # the phase term is rounded to float32 first, as torchaudio's is
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kern = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * window * (base / orig)
    return kern.to(torch.float32).contiguous(), width


class Resample(nn.Module):
    """``torchaudio.transforms.Resample(orig_freq, new_freq)`` as used for ``utils.R1`` / ``utils.R2`` (utils.py:48-49):
    sinc interpolation with a Hann window, ``lowpass_filter_width=6``, ``rolloff=0.99``.  [..., L] -> [..., ceil(new*L/orig)]."""

    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000, resampling_method: str = "sinc_interp_hann",
                 lowpass_filter_width: int = 6, rolloff: float = 0.99, beta=None, *, dtype=None) -> None:
        super().__init__()
        if resampling_method not in ("sinc_interp_hann", "sinc_interpolation"):
            raise NotImplementedError("Resample: only the sinc_interp_hann method is implemented")
        if int(orig_freq) != orig_freq or int(new_freq) != new_freq:
            raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
        import math

        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self.gcd = math.gcd(self.orig_freq, self.new_freq)
        self.lowpass_filter_width, self.rolloff = lowpass_filter_width, rolloff
        self._orig, self._new = self.orig_freq // self.gcd, self.new_freq // self.gcd
        self._table, self._width = (None, 0) if self.orig_freq == self.new_freq else _sinc_hann_table(self._orig, self._new, lowpass_filter_width, rolloff)
        self._native = {}

    def _handle(self, device):
        import ctypes as C
        import weakref

        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._native:
            h = C.c_void_p()
            with torch.cuda.device(device):
                _cabi.check(_cabi.lib().b2d_resampler_create(self._orig, self._new, self._table.shape[1], self._width, self._table.data_ptr(), C.byref(h)))
            weakref.finalize(self, _cabi.lib().b2d_resampler_destroy, h)
            self._native[idx] = h
        return self._native[idx]

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        x = require_cuda_f32(waveform, "waveform")
        if self.orig_freq == self.new_freq:
            return x
        x, lead = _pack(x, 1)
        B, L = x.shape
        lib = _cabi.lib()
        h = self._handle(x.device)
        out = torch.empty((B, lib.b2d_resample_length(h, L)), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(lib.b2d_resample(h, x.data_ptr(), B, L, out.data_ptr(), stream_ptr(x.device)))
        return out.reshape(lead + out.shape[-1:])


def pcm16_to_float(pcm: torch.Tensor, channel: int = 0) -> torch.Tensor:
    """int16 CUDA tensor [n, channels] (or [n]) -> float32 [n]: ``pcm[:, channel] / 32767`` (app3.py:168-172); ``channel=-1``
    averages the channels (app.py:184-186)."""
    if not pcm.is_cuda or pcm.dtype != torch.int16:
        raise TypeError("pcm16_to_float expects an int16 CUDA tensor")
    p = pcm.contiguous()
    n, ch = (p.shape[0], 1) if p.dim() == 1 else (p.shape[0], p.shape[1])
    out = torch.empty(n, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.lib().b2d_pcm16_to_float(p.data_ptr(), n, ch, channel if ch > 1 else 0, out.data_ptr(), stream_ptr(p.device)))
    return out


def float_to_pcm16(x: torch.Tensor) -> torch.Tensor:
    """``(clip(x, -1, 1) * 32767).astype(int16)`` (app3.py:244-245) on the device."""
    x = require_cuda_f32(x, "x")
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().b2d_float_to_pcm16(x.data_ptr(), x.numel(), out.data_ptr(), stream_ptr(x.device)))
    return out
