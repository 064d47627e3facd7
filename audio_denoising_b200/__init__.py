"""audio_denoising_b200 -- B200 (sm_100a) implementation of the belacks/audio-denoising inference hot path.

waveform -> STFT -> Mel log-magnitude -> GRUUNet2 -> inverse Mel -> Griffin-Lim -> iSTFT/OLA, as hand-written
CUDA kernels in ``libb200denoise.so`` (C-ABI: ``include/b200denoise.h``) behind the reference's own Python
interfaces: ``GRUUNet2`` (gruunet2.py), the five torchaudio-style transforms the reference constructs
(app3.py:135-153), and the ``utils`` names.  CUDA only; the package raises if the library is missing.
"""
from . import _cabi
from .checkpoint import TrainingContext, load_denoising_model, save_checkpoint
from .gruunet2 import GRUUNet, GRUUNet2
from .momo3 import MOMO3
from .pipeline import DenoisePipeline, StreamingDenoiser
from .serving import DenoiseServer
from .transforms import GriffinLim, InverseMelScale, InverseSpectrogram, MelScale, Resample, Spectrogram, float_to_pcm16, pcm16_to_float

__all__ = [
    "GRUUNet2", "GRUUNet", "MOMO3", "Spectrogram", "MelScale", "InverseMelScale", "GriffinLim", "InverseSpectrogram", "Resample", "pcm16_to_float", "float_to_pcm16",
    "DenoisePipeline", "StreamingDenoiser", "DenoiseServer", "TrainingContext", "load_denoising_model", "save_checkpoint",
    "native_library",
]
__version__ = "0.1.0"


def native_library():
    """Load libb200denoise.so (raises ImportError when it has not been built)."""
    return _cabi.lib()
