"""Host-side plumbing shared by the drop-in modules: tensor checks, native plan cache, workspaces.

PyTorch is used here for device memory, streams and RNG only -- every computation on the path is a
kernel of libb200denoise.so reached through ``_cabi``.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
import warnings
import weakref

import torch

from . import _cabi


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: audio_denoising_b200 runs on CUDA (sm_100a) only and has no CPU fallback"
        )
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def require_cuda_c64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: audio_denoising_b200 runs on CUDA (sm_100a) only and has no CPU fallback"
        )
    if t.dtype != torch.complex64:
        raise TypeError(f"{name} must be complex64, got {t.dtype}")
    return t.contiguous()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# ---- filterbank (host, float32 -- same arithmetic order as TA:functional/functional.py:518-588) ----
def melscale_fbanks_htk(n_freqs: int, n_mels: int, sample_rate: int) -> torch.Tensor:
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + float(sample_rate // 2) / 700.0)
    m_pts = torch.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    gaps = f_pts[1:] - f_pts[:-1]
    rel = f_pts[None, :] - all_freqs[:, None]
    rising = (-1.0 * rel[:, :-2]) / gaps[:-1]
    falling = rel[:, 2:] / gaps[1:]
    return torch.clamp(torch.minimum(rising, falling), min=0.0).contiguous()


# plan flags (include/b200denoise.h B2D_PLAN_*): fixed at plan creation, nothing on the compute path reads the environment
PLAN_GENERIC_KERNELS, PLAN_EXACT_SQRT, PLAN_EXACT_UNIT, PLAN_EXACT_PEAK_DIV, PLAN_FP32_INVMEL = 1, 2, 4, 8, 16
PLAN_CLUSTER_GL = 32
PLAN_EXACT_ALL = PLAN_EXACT_SQRT | PLAN_EXACT_UNIT | PLAN_EXACT_PEAK_DIV | PLAN_FP32_INVMEL


class Plan:
    """Native DSP plan for one (n_fft, hop, n_mels, sample_rate, flags) on one device."""

    def __init__(self, n_fft: int, hop: int, n_mels: int, sample_rate: int, device: torch.device, flags: int = 0):
        self.n_fft, self.hop, self.n_mels, self.sample_rate, self.flags = n_fft, hop, n_mels, sample_rate, flags
        self.device = device
        self.n_freqs = n_fft // 2 + 1
        if n_mels == 0:  # STFT-only plan (Spectrogram / GriffinLim / InverseSpectrogram): dummy 1-column filterbank
            fb = torch.zeros(self.n_freqs, 1)
            pinv = torch.zeros(self.n_freqs, 1)
            self.rank = 0
        else:
            fb = melscale_fbanks_htk(self.n_freqs, n_mels, sample_rate)
            if (fb.max(dim=0).values == 0.0).any():
                warnings.warn(
                    "At least one mel filterbank has all zero values. "
                    f"The value for `n_mels` ({n_mels}) may be set too high. "
                    f"Or, the value for `n_freqs` ({self.n_freqs}) may be set too low."
                )
            fbt = fb.double().t()  # [n_mels, F]
            self.rank = int(torch.linalg.matrix_rank(fbt))
            pinv = torch.linalg.pinv(fbt).float().contiguous()  # [F, n_mels]
        self.fb = fb
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _cabi.check(_cabi.lib().b2d_plan_create_ex(n_fft, hop, max(n_mels, 1), fb.data_ptr(), pinv.data_ptr(), int(flags), C.byref(handle)))
        self.handle = handle
        self.frame_stride = _cabi.lib().b2d_plan_frame_stride(handle)
        weakref.finalize(self, _cabi.lib().b2d_plan_destroy, handle)

    def num_frames(self, length: int) -> int:
        return 1 + length // self.hop

    def out_length(self, frames: int) -> int:
        return self.hop * (frames - 1)


_plans: dict = {}
_plans_lock = threading.Lock()


def get_plan(n_fft: int, hop: int, n_mels: int, sample_rate: int, device: torch.device, flags: int = 0) -> Plan:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"audio_denoising_b200 plans live on CUDA devices, got {device}")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (n_fft, hop, n_mels, sample_rate, idx, int(flags))
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = Plan(n_fft, hop, n_mels, sample_rate, torch.device("cuda", idx), int(flags))
            _plans[key] = p
    return p


class Workspace:
    """Grow-only byte buffers (caller-owned scratch for the native calls), one per (device, CUDA stream): work enqueued on
    different streams may overlap on the GPU, so it must not share scratch memory."""

    def __init__(self):
        self._bufs: dict = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        device = torch.device(device)
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


def draw_seed() -> int:
    """Non-zero 62-bit seed for the in-kernel Griffin-Lim initial-phase draws, taken from torch's CPU generator."""
    return int(torch.randint(1, 2**62, (1,), dtype=torch.int64).item())


def derive_seed(base: int, index: int) -> int:
    """Non-zero 62-bit seed number ``index`` of a stream keyed by ``base`` (splitmix64 finaliser in Python integers, ~1 us): the
    streaming denoiser draws ``base`` from torch's generator once and derives one seed per hop instead of a ``torch.randint``
    call (~8 us) on every 16 - 20 ms hop."""
    x = (base + 0xD1B54A32D192ED03 * (index + 1)) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return (x & ((1 << 62) - 1)) or 1


def is_hann(window_fn, win_length: int, wkwargs=None) -> bool:
    if window_fn is torch.hann_window and not wkwargs:
        return True
    try:
        w = window_fn(win_length) if wkwargs is None else window_fn(win_length, **wkwargs)
    except Exception:
        return False
    return bool(torch.equal(w.float().cpu(), torch.hann_window(win_length)))
