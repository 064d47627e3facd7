"""Names of the reference's ``utils.py`` kept for drop-in imports (``from utils import *`` in gruunet2.py:17, server.py:20).

None of these is executed on the inference hot path (SURVEY.md L1'): the constants and the small tensor helpers are
plain torch, the PyAV / sounddevice / matplotlib I/O helpers of utils.py:98-398 are out of scope for this build (they
decode files on the host) and raise ``NotImplementedError`` naming the reference function they stand for.
"""
from __future__ import annotations

import torch

SR = 48000  # utils.py:27

# per-frequency-bin standard deviations used by normalize()/denormalize(), utils.py:401-427 (241 bins)
STDS = torch.tensor(
    [0.3922, 0.2043, 0.2245, 0.1914, 0.1832, 0.1889, 0.1823, 0.1581, 0.1304, 0.1081, 0.0921, 0.0825, 0.0775, 0.0758,
    0.0749, 0.0713, 0.0643, 0.0567, 0.0501, 0.0443, 0.0398, 0.0376, 0.0366, 0.0371, 0.0376, 0.0372, 0.0356, 0.0324,
    0.0289, 0.0254, 0.0231, 0.0221, 0.0214, 0.0218, 0.0223, 0.0227, 0.0227, 0.0221, 0.0209, 0.0192, 0.0173, 0.0159,
    0.015, 0.0141, 0.013, 0.0123, 0.0119, 0.0112, 0.0107, 0.0101, 0.0098, 0.0097, 0.0095, 0.0095, 0.0097, 0.0096,
    0.0098, 0.0099, 0.0096, 0.0094, 0.0092, 0.009, 0.0088, 0.0086, 0.0084, 0.0081, 0.0079, 0.0077, 0.0075, 0.0073,
    0.0072, 0.0072, 0.007, 0.0068, 0.0067, 0.0066, 0.0067, 0.0066, 0.0065, 0.0064, 0.0065, 0.0066, 0.0068, 0.0068,
    0.0068, 0.0067, 0.0067, 0.0066, 0.0065, 0.0065, 0.0064, 0.0063, 0.0063, 0.0063, 0.0063, 0.0063, 0.0062, 0.0062,
    0.0061, 0.0062, 0.0062, 0.0062, 0.0061, 0.0061, 0.0062, 0.0062, 0.0063, 0.0062, 0.0062, 0.0061, 0.006, 0.0059,
    0.006, 0.0061, 0.006, 0.0061, 0.0061, 0.0062, 0.0063, 0.0063, 0.0063, 0.0062, 0.0061, 0.0061, 0.0059, 0.0059,
    0.0057, 0.0056, 0.0056, 0.0055, 0.0056, 0.0056, 0.0055, 0.0055, 0.0054, 0.0052, 0.0051, 0.0051, 0.005, 0.0049,
    0.0048, 0.0048, 0.0048, 0.0047, 0.0047, 0.0045, 0.0044, 0.0043, 0.0043, 0.004, 0.0029, 0.0024, 0.0021, 0.0019,
    0.0018, 0.0017, 0.0016, 0.0015, 0.0015, 0.0014, 0.0014, 0.0014, 0.0013, 0.0013, 0.0013, 0.0012, 0.0012, 0.0012,
    0.0012, 0.0012, 0.0011, 0.0011, 0.0011, 0.0011, 0.0011, 0.0011, 0.0011, 0.0011, 0.001, 0.001, 0.001, 0.001, 0.001,
    0.001, 0.001, 0.001, 0.001, 0.001, 0.001, 0.001, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009,
    0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009, 0.0009,
    0.0009, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008,
    0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008, 0.0008,
    0.0008]
)


def _resamplers():
    from .transforms import Resample

    return Resample(44100, SR), Resample(SR, 44100)


R1, R2 = _resamplers()  # utils.py:48-49 (44.1 kHz <-> 48 kHz), running on the B200 polyphase kernel


def normalize(x):
    """utils.py:429-432: divide [B, F, T] (or [B, C, F, T]) spectrogram features by the per-bin STDS."""
    shape = (1, -1, 1) if x.dim() == 3 else (1, 1, -1, 1)
    return x / STDS.to(x.device).view(*shape)


def denormalize(x):
    """utils.py:434-437: inverse of normalize()."""
    shape = (1, -1, 1) if x.dim() == 3 else (1, 1, -1, 1)
    return x * STDS.to(x.device).view(*shape)


def clamp(x):
    """utils.py:82-88: signed log compression sign(x) * log(1 + |x|)."""
    return torch.log(x.abs() + 1) * torch.sign(x)


def unclamp(y):
    """utils.py:89-95: inverse of clamp()."""
    return torch.sign(y) * (torch.exp(y.abs()) - 1)


def unwrap_complex(z):
    """utils.py:70-72: complex [B, ...] -> real [B, 2, ...] (real, imag stacked on axis 1)."""
    return torch.stack([z.real, z.imag]).transpose(0, 1)


def wrap_complex(x, device=None):
    """utils.py:74-80: real [B, 2, ...] -> complex64 [B, ...]."""
    xt = x.transpose(0, 1)
    out = torch.complex(xt[0].float(), xt[1].float())
    return out if device is None else out.to(device)


def _host_io(name):
    def stub(*args, **kwargs):
        raise NotImplementedError(
            f"utils.{name} is a host-side audio I/O helper of the reference (PyAV / sounddevice / matplotlib); "
            "it is outside the GPU hot path this package implements"
        )

    stub.__name__ = name
    return stub


for _n in ['get_canonical_filename', 'figsize_as', 'get_random_audio_buffer', 'collect_random_audio_until_meets_buffer', 'stream_random_audio_buffer', 'plot', 'imshow', 'read_audio', 'play_audio', 'buffer_stream', '__stream_audio', 'stream_audio', 'limit_stream', '__combine_samples', 'combine_audio', 'clip_audio_to_same_size']:
    globals()[_n] = _host_io(_n)
del _n
