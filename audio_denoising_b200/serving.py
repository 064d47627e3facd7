"""The request loop of the reference's socket inference server (server.py:166-227) on the B200 path
(SURVEY.md section 8f rank 1).

Wire protocol kept as is: a ``multiprocessing.connection.Listener`` on ``('localhost', 6101)``; each message is a pickled
``numpy`` array ``[n_samples, n_channels]``; channel 0 is denoised with the noisy-phase chain (server.py:207-216:
STFT -> Mel log-magnitude -> GRUUNet2 -> ``relu(pred) * 3`` -> inverse Mel of ``exp(logmel - out) - 1`` -> iSTFT with the
noisy phase), the GRU state persists across requests and connections with the ``hx *= 0.9`` leak (server.py:177,214), and
the answer is the denoised signal repeated over the input's channel count, ``[hop * (T - 1), n_channels]``.  The string
``'close'`` closes the connection; errors restart the listener after 0.1 s (server.py:224-227).
"""
from __future__ import annotations

import time
from multiprocessing.connection import Listener
from typing import Optional

import numpy as np
import torch

from .gruunet2 import GRUUNet2
from .pipeline import DenoisePipeline


class DenoiseServer:
    def __init__(self, model: GRUUNet2, n_fft: int = 1024, hop_length: int = 512, n_mels: int = 64, sample_rate: int = 48000,
                 out_scale: float = 3.0, hx_decay: float = 0.9, device: Optional[torch.device] = None):
        self.pipe = DenoisePipeline(model, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels, sample_rate=sample_rate, device=device)
        self.out_scale, self.hx_decay = out_scale, hx_decay
        self.hx: Optional[torch.Tensor] = None  # module-global ``hx`` of server.py:177
        self.device = self.pipe.device

    @torch.no_grad()
    def handle(self, X: np.ndarray) -> np.ndarray:
        """One request: X [n, channels] (or [n]) -> [hop*(T-1), channels] float32 (server.py:199-220)."""
        x = torch.as_tensor(np.asarray(X), dtype=torch.float32).T  # server.py:199
        n_channels = 1
        if x.dim() == 2:
            n_channels = x.shape[0]
            x = x[0].reshape(1, -1)  # "monotize", server.py:201-205
        else:
            x = x.reshape(1, -1)
        wave, self.hx = self.pipe.denoise_noisy_phase(x.to(self.device), self.hx, self.out_scale, self.hx_decay)
        out = wave.repeat(n_channels, 1)  # server.py:216
        return out.T.cpu().numpy()

    def serve_forever(self, address=("localhost", 6101), accept_timeout: float = 5.0, max_requests: Optional[int] = None) -> int:
        """Accept loop of server.py:181-227.  Returns the number of requests served (``max_requests`` bounds it for tests)."""
        served = 0
        while max_requests is None or served < max_requests:
            try:
                with Listener(address) as listener:
                    listener._listener._socket.settimeout(accept_timeout)  # server.py:184
                    while max_requests is None or served < max_requests:
                        with listener.accept() as conn:
                            while max_requests is None or served < max_requests:
                                try:
                                    X = conn.recv()
                                except Exception:  # peer went away: back to accept (server.py:193-196)
                                    break
                                if isinstance(X, str) and X == "close":
                                    break
                                conn.send(self.handle(X))
                                served += 1
            except KeyboardInterrupt:
                break
            except Exception:
                time.sleep(0.1)  # server.py:224-227
        return served
