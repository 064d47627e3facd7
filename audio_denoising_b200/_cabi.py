"""ctypes binding of libb200denoise.so (C-ABI declared in include/b200denoise.h).

This is the only place the Python host side touches the native library.  It fails loudly:
a missing / unloadable library raises ``ImportError`` at first use, a non-zero status raises
``B2DError`` carrying ``b2d_last_error_string()``.  There is no CPU fallback anywhere.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200denoise.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_ALIGN, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4, -5


class B2DError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200denoise error {code}: {msg}")
        self.code = code


class ModelConfig(C.Structure):
    _fields_ = [
        ("num_compressed_bins", C.c_int),
        ("hidden", C.c_int),
        ("levels", C.c_int),
        ("kernel", C.c_int),
        ("stride", C.c_int),
        ("padding", C.c_int),
        ("num_gaussians", C.c_int),
    ]


class CellConfig(C.Structure):
    _fields_ = [
        ("arch", C.c_int),
        ("num_compressed_bins", C.c_int),
        ("levels", C.c_int),
        ("num_gaussians", C.c_int),
        ("n_mels", C.c_int),
        ("hidden", C.c_int * 8),
        ("kernel", C.c_int * 8),
        ("stride", C.c_int * 8),
        ("padding", C.c_int * 8),
    ]


_vp, _i, _f, _sz, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_ulonglong

# name -> (restype, argtypes); mirrors include/b200denoise.h one to one
SIGNATURES = {
    "b2d_version": (_i, []),
    "b2d_last_error_string": (C.c_char_p, []),
    "b2d_launch_count": (C.c_ulonglong, []),
    "b2d_plan_create": (_i, [_i, _i, _i, _vp, _vp, C.POINTER(_vp)]),
    "b2d_plan_create_ex": (_i, [_i, _i, _i, _vp, _vp, C.c_uint, C.POINTER(_vp)]),
    "b2d_plan_destroy": (None, [_vp]),
    "b2d_plan_num_frames": (_i, [_vp, _i]),
    "b2d_plan_output_length": (_i, [_vp, _i]),
    "b2d_plan_frame_stride": (_i, [_vp]),
    "b2d_model_create": (_i, [C.POINTER(ModelConfig), C.POINTER(_vp), _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "b2d_model_destroy": (None, [_vp]),
    "b2d_model_n_mels": (_i, [_vp]),
    "b2d_cell_create": (_i, [C.POINTER(CellConfig), C.POINTER(_vp), _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "b2d_cell_destroy": (None, [_vp]),
    "b2d_cell_workspace_bytes": (_sz, [_vp, _i, _i]),
    "b2d_cell_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "b2d_cell_num_param_floats": (_i, [_vp]),
    "b2d_cell_backward_workspace_bytes": (_sz, [_vp, _i, _i]),
    "b2d_cell_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "b2d_peak": (_i, [_vp, _i, _i, _vp, _vp]),
    "b2d_stft": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b2d_stft_mel_log1p": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "b2d_mel_scale": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b2d_gruunet2_workspace_bytes": (_sz, [_vp, _i, _i]),
    "b2d_gruunet2_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "b2d_residual_mel": (_i, [_vp, _vp, _vp, _sz, _i, _f, _vp]),
    "b2d_inverse_mel": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b2d_inverse_mel_frames": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b2d_griffinlim_workspace_bytes": (_sz, [_vp, _i, _i]),
    "b2d_griffinlim": (_i, [_vp, _vp, _vp, _u64, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "b2d_griffinlim_frames": (_i, [_vp, _vp, _vp, _u64, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "b2d_istft": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b2d_denoise_workspace_bytes": (_sz, [_vp, _vp, _i, _i]),
    "b2d_denoise_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _u64, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2d_denoise_pcm16_workspace_bytes": (_sz, [_vp, _vp, _i, _i]),
    "b2d_denoise_batch_pcm16": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _u64, _i, _f, _i, _i, _vp, _vp, _sz, _vp]),
    "b2d_denoise_noisy_phase_workspace_bytes": (_sz, [_vp, _vp, _i, _i]),
    "b2d_denoise_noisy_phase": (_i, [_vp, _vp, _vp, _i, _i, _vp, _f, _f, _i, _vp, _vp, _sz, _vp]),
    "b2d_stream_step_workspace_bytes": (_sz, [_vp, _vp, _i]),
    "b2d_stream_step": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _u64, _vp, _i, _f, _i, _vp, _vp, _sz, _vp]),
    "b2d_pcm16_to_float": (_i, [_vp, _sz, _i, _i, _vp, _vp]),
    "b2d_float_to_pcm16": (_i, [_vp, _sz, _vp, _vp]),
    "b2d_resampler_create": (_i, [_i, _i, _i, _i, _vp, C.POINTER(_vp)]),
    "b2d_resampler_destroy": (None, [_vp]),
    "b2d_resample_length": (_i, [_vp, _i]),
    "b2d_resample": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load (once) and return the native library; raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -m audio_denoising_b200._build` "
                    "(nvcc, sm_100a). audio_denoising_b200 has no CPU / PyTorch fallback."
                )
            try:
                handle = C.CDLL(LIB_PATH)
            except OSError as e:  # pragma: no cover
                raise ImportError(f"cannot load {LIB_PATH}: {e}") from e
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise B2DError(rc, lib().b2d_last_error_string().decode(errors="replace"))


def launch_count() -> int:
    return int(lib().b2d_launch_count())
