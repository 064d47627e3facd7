"""Whole-chain entry points: the inline glue of app3.py:123-250 and server.py:166-220 as objects.

* ``DenoisePipeline.denoise``            -- app3.py:181-217 applied to whole clips [B, L]
* ``DenoisePipeline.denoise_noisy_phase``-- server.py:207-216 (noisy-phase iSTFT, ``hx *= 0.9`` leak)
* ``DenoisePipeline.denoise_host``       -- same as ``denoise`` from / to pinned host memory, copies overlapped
* ``StreamingDenoiser``                  -- ``DenoisingAudioProcessor`` hop loop (app3.py:167-226) with device-resident state

Each call is ONE native call (``b2d_denoise_batch`` / ``b2d_denoise_noisy_phase`` / ``b2d_stream_step``)
that enqueues the kernel chain on the current CUDA stream.
"""
from __future__ import annotations

import threading
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._runtime import Workspace, derive_seed, draw_seed, get_plan, ptr, require_cuda_c64, require_cuda_f32, stream_ptr
from .gruunet2 import CONV_MODES, GRUUNet2


def _indexed_device(device) -> torch.device:
    """``cuda`` / None -> the current device with an explicit index, so device comparisons are exact."""
    d = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if d.type != "cuda":
        raise RuntimeError(f"audio_denoising_b200 runs on CUDA devices only, got {d}")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


class DenoisePipeline:
    def __init__(self, model: GRUUNet2, n_fft: int = 1024, hop_length: int = 512, n_mels: int = 64, sample_rate: int = 16000,
                 n_iter: int = 32, momentum: float = 0.99, device: Optional[torch.device] = None, plan_flags: int = 0):
        if not isinstance(model, GRUUNet2):
            raise TypeError("model must be an audio_denoising_b200.GRUUNet2")
        if not model.uses_tuned_kernels():
            raise NotImplementedError("DenoisePipeline fuses the shipped GRUUNet2 configuration (hidden 17 x 4, 4 bins, k3 s2 p1); "
                                      "other configurations run through model(x) and the transform modules")
        if n_mels != model.n_mels:
            raise ValueError(f"n_mels={n_mels} does not match the model ({model.n_mels})")
        if hop_length * 2 != n_fft:
            raise NotImplementedError("DenoisePipeline: only hop_length == n_fft // 2 is implemented (the reference's setting)")
        self.model = model
        self.n_fft, self.hop, self.n_mels, self.sample_rate = n_fft, hop_length, n_mels, sample_rate
        self.n_iter, self.momentum = n_iter, momentum
        self.device = _indexed_device(device)
        self.plan = get_plan(n_fft, hop_length, n_mels, sample_rate, self.device, plan_flags)  # flags: _runtime.PLAN_*
        # device staging slots of denoise_host (set before the first call).  Three instead of two lets the upload of batch i + 1 start
        # two batches ahead: worth 3 % where eight GPUs share one host (4.21 -> 4.10 ms per step, tools/e2e_probe.py), nothing at N = 1
        self.host_ring_depth = 3
        if self.plan.rank < n_mels:
            raise ValueError(f"mel filterbank is rank deficient ({self.plan.rank} < {n_mels}) for n_fft={n_fft}, sample_rate={sample_rate}")
        self._ws = Workspace()  # one scratch buffer per CUDA stream the pipeline is used on
        self._lock = threading.Lock()  # a call's launches are enqueued as one unit (threads sharing a stream share its scratch)
        self._host = None

    # -- shapes ---------------------------------------------------------------------------------
    def num_frames(self, L: int) -> int:
        return self.plan.num_frames(L)

    def out_length(self, L: int) -> int:
        return self.plan.out_length(self.plan.num_frames(L))

    def _hx(self, hx, B, device):
        shape = (B, self.model.latent_size, self.model.num_compressed_bins)
        if hx is None:
            return torch.zeros(shape, dtype=torch.float32, device=device)
        h = require_cuda_f32(hx, "hx")
        if tuple(h.shape) != shape:
            raise ValueError(f"hx must be {list(shape)} (batch, latent, compressed bins), got {list(h.shape)}")
        if h.device != device:
            raise ValueError(f"hx is on {h.device}, the input on {device}")
        return h.clone()  # the reference never mutates the caller's hx

    def _check_input(self, x: torch.Tensor, name: str):
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.dim() != 2:
            raise ValueError(f"{name} must be [B, L] (or [L]), got {list(x.shape)}")
        if x.device != self.device:
            raise ValueError(f"{name} is on {x.device} but this pipeline (plan, packed model) lives on {self.device}")
        return x

    def _check_out(self, out, B, Lout, dtype, device):
        if out is None:
            return torch.empty((B, Lout), dtype=dtype, device=device)
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.device == device and out.dtype == dtype
                and tuple(out.shape) == (B, Lout) and out.is_contiguous()):
            raise ValueError(f"out must be a contiguous {dtype} tensor of shape [{B}, {Lout}] on {device}")
        return out

    # -- app3 chain -----------------------------------------------------------------------------
    @torch.no_grad()
    def denoise(self, noisy: torch.Tensor, hx: Optional[torch.Tensor] = None, init_angles: Optional[torch.Tensor] = None,
                normalise: bool = True, rand_init: bool = True, return_intermediates: bool = False, out: Optional[torch.Tensor] = None):
        """noisy [B, L] (CUDA float32) -> wave [B, hop*(T-1)] ; returns (wave, hx) or a dict."""
        x = self._check_input(require_cuda_f32(noisy, "noisy"), "noisy")
        B, L = x.shape
        T = self.plan.num_frames(L)
        F = self.plan.n_freqs
        dev = x.device
        h = self._hx(hx, B, dev)
        seed = draw_seed() if (init_angles is None and rand_init) else 0  # TA functional.py:310, drawn in-kernel
        if init_angles is not None:
            init_angles = require_cuda_c64(init_angles, "init_angles").reshape(B, F, T)
        wave = self._check_out(out, B, self.plan.out_length(T), torch.float32, dev)
        logmel = pred = mag = None
        if return_intermediates:
            logmel = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=dev)
            pred = torch.empty_like(logmel)
            mag = torch.empty((B, T, self.plan.frame_stride), dtype=torch.float32, device=dev)
        lib = _cabi.lib()
        native = self.model.native_model(dev)
        with self._lock, torch.cuda.device(dev):
            ws = self._ws.get(lib.b2d_denoise_workspace_bytes(self.plan.handle, native.handle, B, L), dev)
            _cabi.check(lib.b2d_denoise_batch(
                self.plan.handle, native.handle, x.data_ptr(), B, L, h.data_ptr(), ptr(init_angles), seed, self.n_iter, float(self.momentum),
                1 if normalise else 0, self.model.native_conv_mode(), wave.data_ptr(), ptr(logmel), ptr(pred), ptr(mag),
                ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        if return_intermediates:
            return dict(wave=wave, hx=h, logmel=logmel.transpose(-1, -2), pred=pred, lin_mag=mag[..., :F].transpose(-1, -2))
        return wave, h

    @torch.no_grad()
    def denoise_pcm16(self, pcm: torch.Tensor, hx: Optional[torch.Tensor] = None, init_angles: Optional[torch.Tensor] = None,
                      normalise: bool = True, rand_init: bool = True, out: Optional[torch.Tensor] = None):
        """The same chain on the int16 PCM link of ``recv`` (app3.py:168-172 in, :244-245 out): pcm [B, L] int16 CUDA ->
        (pcm_out [B, hop*(T-1)] int16, hx).  One native call; ``x = pcm / 32767`` is fused with the peak pass and the
        float staging lives in the call's workspace -- nothing is allocated per call once the workspace exists."""
        if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda and pcm.dtype == torch.int16):
            raise TypeError("denoise_pcm16 expects an int16 CUDA tensor")
        x = self._check_input(pcm.contiguous(), "pcm")
        B, L = x.shape
        T = self.plan.num_frames(L)
        dev = x.device
        h = self._hx(hx, B, dev)
        seed = draw_seed() if (init_angles is None and rand_init) else 0
        if init_angles is not None:
            init_angles = require_cuda_c64(init_angles, "init_angles").reshape(B, self.plan.n_freqs, T)
        res = self._check_out(out, B, self.plan.out_length(T), torch.int16, dev)
        lib = _cabi.lib()
        native = self.model.native_model(dev)
        with self._lock, torch.cuda.device(dev):
            ws = self._ws.get(lib.b2d_denoise_pcm16_workspace_bytes(self.plan.handle, native.handle, B, L), dev)
            _cabi.check(lib.b2d_denoise_batch_pcm16(
                self.plan.handle, native.handle, x.data_ptr(), B, L, h.data_ptr(), ptr(init_angles), seed, self.n_iter, float(self.momentum),
                1 if normalise else 0, self.model.native_conv_mode(), res.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        return res, h

    # -- server chain -----------------------------------------------------------------------------
    @torch.no_grad()
    def denoise_noisy_phase(self, x: torch.Tensor, hx: Optional[torch.Tensor] = None, out_scale: float = 3.0, hx_decay: float = 0.9):
        """server.py:207-216: x [B, L] -> (wave [B, hop*(T-1)], hx) using the noisy phase."""
        x = self._check_input(require_cuda_f32(x, "x"), "x")
        B, L = x.shape
        T = self.plan.num_frames(L)
        dev = x.device
        h = self._hx(hx, B, dev)
        wave = torch.empty((B, self.plan.out_length(T)), dtype=torch.float32, device=dev)
        lib = _cabi.lib()
        native = self.model.native_model(dev)
        with self._lock, torch.cuda.device(dev):
            ws = self._ws.get(lib.b2d_denoise_noisy_phase_workspace_bytes(self.plan.handle, native.handle, B, L), dev)
            _cabi.check(lib.b2d_denoise_noisy_phase(
                self.plan.handle, native.handle, x.data_ptr(), B, L, h.data_ptr(), float(out_scale), float(hx_decay),
                self.model.native_conv_mode(), wave.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        return wave, h

    # -- host-to-host (end-to-end) ----------------------------------------------------------------
    @torch.no_grad()
    def denoise_host(self, noisy_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, chunks: int = 1,
                     rand_init: bool = True, init_angles: Optional[torch.Tensor] = None, wait: bool = True) -> torch.Tensor:
        """noisy_host [B, L] pinned CPU float32 -> pinned CPU [B, hop*(T-1)].

        Copies run on their own streams: host->device on ``h2d``, kernels on the current stream, device->host on
        ``d2h``, ordered by events.  With ``wait=False`` the call only enqueues work, so consecutive batches pipeline
        (batch i+1's upload and batch i-1's download overlap batch i's kernels); call ``host_synchronize()`` before
        reading ``out_host``.  ``chunks > 1`` additionally slices one batch so its own copies overlap its own kernels.
        """
        if noisy_host.is_cuda:
            raise ValueError("denoise_host takes host tensors; use denoise() for device tensors")
        pcm = noisy_host.dtype == torch.int16  # int16 PCM in -> int16 PCM out (app3.py:168-172 / 244-245), half the link bytes
        if not pcm and noisy_host.dtype != torch.float32:
            raise TypeError("denoise_host takes float32 waveforms or int16 PCM")
        B, L = noisy_host.shape
        T = self.plan.num_frames(L)
        Lout = self.plan.out_length(T)
        if out_host is None:
            out_host = torch.empty((B, Lout), dtype=noisy_host.dtype, pin_memory=True)
        elif out_host.dtype != noisy_host.dtype:
            raise TypeError("out_host must have the dtype of noisy_host")
        dev = self.device
        chunks = max(1, min(chunks, B))
        bounds = [(i * B) // chunks for i in range(chunks + 1)]
        if self._host is None:
            self._host = dict(h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev), slots={}, turn={})
        h2d, d2h = self._host["h2d"], self._host["d2h"]
        compute = torch.cuda.current_stream(dev)
        for i in range(chunks):
            lo, hi = bounds[i], bounds[i + 1]
            # Device staging is allocated once per shape and double-buffered (slot = two tensors + the events that say when
            # they are free again): nothing is allocated per call, so the caching allocator never has to synchronise streams.
            key = (hi - lo, L, noisy_host.dtype)
            ring = self._host["slots"].get(key)
            if ring is None:
                # xin / res carry the link format (float32, or int16 PCM: half the bytes on the link); for int16 the float
                # staging sits in the native call's workspace
                ring = [dict(xin=torch.empty((hi - lo, L), dtype=noisy_host.dtype, device=dev),
                             res=torch.empty((hi - lo, Lout), dtype=noisy_host.dtype, device=dev),
                             in_free=None, out_free=None) for _ in range(self.host_ring_depth)]
                self._host["slots"][key] = ring
                self._host["turn"][key] = 0
            slot = ring[self._host["turn"][key] % len(ring)]
            self._host["turn"][key] += 1
            with torch.cuda.stream(h2d):
                if slot["in_free"] is not None:
                    h2d.wait_event(slot["in_free"])  # the kernels of two batches ago have consumed this input buffer
                slot["xin"].copy_(noisy_host[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(h2d)
            compute.wait_event(ev_in)
            if slot["out_free"] is not None:
                compute.wait_event(slot["out_free"])  # ... and its result has left for the host
            ia = None if init_angles is None else init_angles[lo:hi]
            if pcm:
                self.denoise_pcm16(slot["xin"], init_angles=ia, rand_init=rand_init, out=slot["res"])
            else:
                self.denoise(slot["xin"], init_angles=ia, rand_init=rand_init, out=slot["res"])
            slot["in_free"] = torch.cuda.Event()
            slot["in_free"].record(compute)
            with torch.cuda.stream(d2h):
                d2h.wait_event(slot["in_free"])
                out_host[lo:hi].copy_(slot["res"], non_blocking=True)
                slot["out_free"] = torch.cuda.Event()
                slot["out_free"].record(d2h)
        if wait:
            self.host_synchronize()
        return out_host

    def host_synchronize(self) -> None:
        """Block until every batch enqueued by ``denoise_host(wait=False)`` has landed in host memory."""
        if self._host is not None:
            self._host["d2h"].synchronize()
        torch.cuda.current_stream(self.device).synchronize()


class StreamingDenoiser:
    """Device-resident equivalent of ``DenoisingAudioProcessor`` (app3.py:123-250) for S parallel sessions.

    ``push(chunk)`` mirrors ``recv``: it appends samples to the input ring and, for every full ``n_fft`` window,
    runs one hop (peak normalise, Hann pre-window, 3-frame STFT -> Mel -> GRUUNet2 (hx carried) -> inverse Mel ->
    Griffin-Lim -> x peak -> overlap-add ring) and returns the ``hop`` samples emitted, one hop late, exactly as
    app3.py:219-226 does (quirks Q2-Q4 kept).  ``angles_fn(hop_index, shape)`` may supply the Griffin-Lim init.
    """

    def __init__(self, model: GRUUNet2, n_fft: int = 1536, hop_length: int = 768, n_mels: int = 64, sample_rate: int = 48000,
                 n_iter: int = 32, momentum: float = 0.99, sessions: int = 1, device: Optional[torch.device] = None, angles_fn=None,
                 use_graph: bool = True):
        if hop_length * 2 != n_fft:
            raise NotImplementedError("StreamingDenoiser: only hop_length == n_fft // 2 is implemented")
        self.model = model
        self.n_fft, self.hop, self.n_mels, self.sample_rate = n_fft, hop_length, n_mels, sample_rate
        self.n_iter, self.momentum, self.S = n_iter, momentum, sessions
        self.device = _indexed_device(device)
        self.plan = get_plan(n_fft, hop_length, n_mels, sample_rate, self.device)
        if self.plan.rank < n_mels:
            raise ValueError(f"mel filterbank is rank deficient ({self.plan.rank} < {n_mels}) for n_fft={n_fft}, sample_rate={sample_rate}")
        self.angles_fn = angles_fn
        self.hx = torch.zeros(sessions, model.latent_size, model.num_compressed_bins, dtype=torch.float32, device=self.device)
        self.ola = torch.zeros(sessions, n_fft, dtype=torch.float32, device=self.device)
        self.inbuf = np.zeros((sessions, 0), dtype=np.float32)
        self.hops = 0
        self._ws = Workspace()
        self._chunk_host = torch.empty((sessions, n_fft), dtype=torch.float32, pin_memory=True)
        self._out_host = torch.empty((sessions, hop_length), dtype=torch.float32, pin_memory=True)
        self._chunk_dev = torch.empty((sessions, n_fft), dtype=torch.float32, device=self.device)
        self._out_dev = torch.empty((sessions, hop_length), dtype=torch.float32, device=self.device)
        self._seed_host = torch.ones(1, dtype=torch.int64).pin_memory()
        self._seed_base = draw_seed()  # from torch's generator, once; hop k uses derive_seed(base, k)
        self._seed_np = self._seed_host.numpy()  # writing through the numpy view costs 0.2 us, tensor.__setitem__ 5 us
        self._chunk_np = self._chunk_host.numpy()
        self._out_np = self._out_host.numpy()
        self._seed_dev = torch.ones(1, dtype=torch.int64, device=self.device)
        self.use_graph = use_graph
        self._graph = None
        self._side = None
        self._graph_native = None  # the NativeModel whose device pointers the captured graph holds (kept alive with it)

    def _native_step(self, init, seed, zero_copy: bool = False):
        """``zero_copy``: the kernels read the chunk and the seed from, and write the emitted hop to, the PINNED HOST buffers
        directly (pinned memory is device-accessible under unified addressing): a few KB per hop cross PCIe inside the first /
        last kernel instead of through three copy nodes of the graph (~2.5 us each)."""
        lib = _cabi.lib()
        dev = self.device
        handle = self.model.native_model(dev).handle
        ws = self._ws.get(lib.b2d_stream_step_workspace_bytes(self.plan.handle, handle, self.S), dev)
        chunk, seed_buf, out = ((self._chunk_host, self._seed_host, self._out_host) if zero_copy
                                else (self._chunk_dev, self._seed_dev, self._out_dev))
        with torch.cuda.device(dev):
            _cabi.check(lib.b2d_stream_step(
                self.plan.handle, handle, chunk.data_ptr(), self.S, self.hx.data_ptr(), self.ola.data_ptr(), ptr(init), seed,
                seed_buf.data_ptr(), self.n_iter, float(self.momentum), self.model.native_conv_mode(), out.data_ptr(),
                ws.data_ptr(), ws.numel(), stream_ptr(dev)))

    def _capture(self):
        """Capture one hop (the whole kernel chain, reading the chunk and the 8-byte seed from and writing the emitted hop to
        pinned host memory) into a CUDA graph: a hop then costs one graph launch instead of ~10 kernel launches + 3 copies."""
        dev = self.device
        native = self.model.native_model(dev)
        state = (self.hx.clone(), self.ola.clone())
        if self._side is None:
            self._side = torch.cuda.Stream(dev)  # capture stream; its workspace entry is reused by a re-capture
        side = self._side
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: module loading, cudaFuncSetAttribute, workspace allocation
                self._native_step(None, 1, zero_copy=True)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            self._native_step(None, 1, zero_copy=True)
        self.hx.copy_(state[0])
        self.ola.copy_(state[1])
        torch.cuda.synchronize(dev)
        self._graph, self._graph_native = graph, native

    @torch.no_grad()
    def step(self, window: np.ndarray) -> np.ndarray:
        """One hop: window [S, n_fft] float32 host -> [S, hop] float32 host (includes H2D and D2H, like app3.py:189,215)."""
        dev = self.device
        self._chunk_np[...] = window
        F = self.plan.n_freqs
        if self.angles_fn is None and self.use_graph:
            # the graph holds the packed model's device pointers: re-capture when the weights were re-packed
            # (load_state_dict / optimizer.step / .to() on a live model); the old pack stays alive until then
            if self._graph is None or self.model.native_model(dev) is not self._graph_native:
                self._capture()
            self._seed_np[0] = derive_seed(self._seed_base, self.hops)
            self._graph.replay()
            torch.cuda.current_stream(dev).synchronize()
        else:
            self._chunk_dev.copy_(self._chunk_host, non_blocking=True)
            init, seed = None, 0
            if self.angles_fn is not None:
                init = require_cuda_c64(self.angles_fn(self.hops, (self.S, F, 3)).to(dev), "init_angles")
            else:
                seed = 1
                self._seed_np[0] = derive_seed(self._seed_base, self.hops)  # the same seed stream as the graph path
                self._seed_dev.copy_(self._seed_host, non_blocking=True)
            self._native_step(init, seed)
            self._out_host.copy_(self._out_dev, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        self.hops += 1
        return self._out_np.copy()

    def push(self, chunk: np.ndarray) -> np.ndarray:
        """chunk: [n] or [S, n] float32 samples (or int16, scaled by 1/32767 as app3.py:172).  Returns [S, k*hop]."""
        c = np.asarray(chunk)
        if c.dtype == np.int16:
            c = c.astype(np.float32) / np.iinfo(np.int16).max
        c = c.astype(np.float32).reshape(self.S, -1) if c.ndim > 1 else np.broadcast_to(c.astype(np.float32), (self.S, c.shape[0]))
        self.inbuf = np.concatenate([self.inbuf, c], axis=1)
        outs = []
        while self.inbuf.shape[1] >= self.n_fft:
            outs.append(self.step(self.inbuf[:, : self.n_fft]))
            self.inbuf = self.inbuf[:, self.hop:]
        if not outs:
            return np.zeros((self.S, 0), dtype=np.float32)
        return np.concatenate(outs, axis=1)

    def recv_pcm(self, pcm: np.ndarray) -> np.ndarray:
        """``DenoisingAudioProcessor.recv`` on raw int16 frames (app3.py:167-250) for a single session: int16 [n] or
        [n, channels] in -> int16 out.  While no hop has been produced yet the input is passed through (padded / cut to the
        frame length), exactly as app3.py:228-241 does."""
        if self.S != 1:
            raise ValueError("recv_pcm mirrors one WebRTC session; use push() for batched sessions")
        a = np.asarray(pcm)
        if a.ndim > 1:
            a = a[:, 0]  # app3.py:169-170
        chunk = a.astype(np.float32) / np.iinfo(np.int16).max
        out = self.push(chunk)
        if out.shape[1] == 0:
            return (np.clip(chunk, -1.0, 1.0) * np.iinfo(np.int16).max).astype(np.int16)
        return self.to_int16(out[0])

    @staticmethod
    def to_int16(x: np.ndarray) -> np.ndarray:
        """app3.py:244-245."""
        return (np.clip(x, -1.0, 1.0) * np.iinfo(np.int16).max).astype(np.int16)
