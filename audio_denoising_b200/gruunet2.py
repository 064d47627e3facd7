"""Drop-in ``GRUUNet2`` (reference: gruunet2.py:246-306) running on the sm_100a kernels.

Same constructor, same ``state_dict`` keys / ``parameters()`` order, same
``forward(input, hx=None) -> (out, hx)`` with the 2-D / 3-D input handling of gruunet2.py:290-306,
same ``.hparams`` / ``.get_config()`` / ``.from_config`` helpers (gruunet2.py:29-51) -- so
``load_state_dict`` of the shipped checkpoints and ``TrainingContext.load`` (server.py:129-142)
work unchanged.  The parameter holders are real ``nn.Conv1d`` / ``nn.ConvTranspose1d`` modules so
that seeded construction draws the same initial weights as the reference, but they are never
*called*: ``forward`` packs the weights into a native model (Gaussian position channels folded into
per-position biases) and runs encoder -> persistent recurrence -> decoder kernels.

Inference (eval() mode or ``torch.no_grad()``, as the reference calls the model: app3.py:200, server.py:211) runs the fused
kernels; in train() mode with autograd recording the call is differentiable through hand-written fp32 backward kernels
(csrc/cell.cu).  CUDA float32 tensors only; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from typing import Optional, Sequence

import torch
from torch import nn

from . import _cabi
from ._runtime import require_cuda_f32, stream_ptr

# convolution engines (include/b200denoise.h B2D_CONV_*): "mma" = warp-level TF32 tensor-core MMAs with the fp32-class big + small
# split (default, fastest); "utc" = the encoder as one persistent tcgen05 / TMEM kernel + the warp-MMA decoder; "fp32" = CUDA-core
# FMA kernels (the exact engine of the attribution test)
CONV_MODES = {"fp32": 0, "mma": 3, "utc": 5}


class _PositionCode(nn.Module):
    """Holder of the ``gs.offset`` buffer (GaussianSmearing centres, gruunet2.py:54-68)."""

    def __init__(self, count: int):
        super().__init__()
        self.register_buffer("offset", torch.linspace(0.0, 1.0, count))


class _Slot(nn.Module):
    """One ``...{i}.conv`` parameter holder."""

    def __init__(self, conv: nn.Module):
        super().__init__()
        self.conv = conv


class _Encoder(nn.Module):  # state_dict prefix "downs.{i}.conv", "gs.offset"
    def __init__(self, channels: Sequence[int], kernels, strides, paddings, count: int):
        super().__init__()
        self.downs = nn.ModuleList()
        self.gs = _PositionCode(count)
        for i in range(len(channels) - 1):
            self.downs.append(_Slot(nn.Conv1d(channels[i] + count, channels[i + 1], kernel_size=kernels[i], stride=strides[i], padding=paddings[i])))


class _Decoder(nn.Module):  # state_dict prefix "ups.{i}.conv", "gs.offset"
    def __init__(self, channels: Sequence[int], kernels, strides, paddings, count: int):
        super().__init__()
        self.ups = nn.ModuleList()
        n = len(channels) - 1
        for i in range(n):
            cin = channels[i] + count if i == 0 else 2 * channels[i] + count
            self.ups.append(_Slot(nn.ConvTranspose1d(cin, channels[i + 1], kernel_size=kernels[i], padding=paddings[i], stride=strides[i])))
        self.gs = _PositionCode(count)


class _Cell(nn.Module):  # "input_gate", "reset_gate", "output_gate"
    def __init__(self, in_size, hidden_sizes, kernel_sizes, strides, paddings, count):
        super().__init__()
        hs = list(hidden_sizes)
        wide = hs[:-1] + [3 * hs[-1]]
        self.input_gate = _Encoder([in_size] + wide, list(kernel_sizes), list(strides), list(paddings), count)
        self.reset_gate = _Encoder([hs[-1], 3 * hs[-1]], [3], [1], [1], count)
        self.output_gate = _Decoder(hs[::-1] + [1], list(kernel_sizes)[::-1], list(strides)[::-1], list(paddings)[::-1], count)


class NativeModel:
    """One packed ``b2d_model`` (weights re-laid-out on one device).  Immutable; destroyed when the last Python reference
    goes away, so a CUDA graph or pipeline that captured its device pointers keeps it alive by holding this object."""

    def __init__(self, handle, signature, device_index: int):
        self.handle, self.signature, self.device_index = handle, signature, device_index
        self._fin = weakref.finalize(self, _cabi.lib().b2d_model_destroy, handle)


class GRUUNet2(nn.Module):
    def __init__(self, num_compressed_bins, in_size, hidden_sizes, kernel_sizes, strides, paddings, num_gaussians=6):
        super().__init__()
        assert in_size == 1
        self.hparams = dict(
            num_compressed_bins=num_compressed_bins, in_size=in_size, hidden_sizes=hidden_sizes, kernel_sizes=kernel_sizes,
            strides=strides, paddings=paddings, num_gaussians=num_gaussians,
        )
        self.latent_size = hidden_sizes[-1]
        self.num_compressed_bins = num_compressed_bins
        self.cell = _Cell(in_size, hidden_sizes, kernel_sizes, strides, paddings, num_gaussians)
        self.conv_mode = "mma"  # one of CONV_MODES
        self.exact_gates = False  # attribution switch: GRU gates through expf / IEEE division (B2D_CONV_EXACT_GATES)
        self._cell_runner = None  # generic cell kernels for configurations other than the shipped one (csrc/cell.cu)
        self._native = {}  # device index -> NativeModel
        self._native_lock = threading.Lock()
        self._generation = 0  # bumped by repack()

    # ---- gruunet2.py:29-51 helpers ------------------------------------------------------------
    def get_config(self):
        return self.hparams

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    @property
    def n_mels(self) -> int:
        """Length of the model's input axis: the conv length formula of gruunet2.py:127-157 run backwards from
        ``num_compressed_bins`` (L_in = (L_out - 1) * s - 2 p + k; ``bins << levels`` for the shipped k3 s2 p1)."""
        hp = self.hparams
        n = self.num_compressed_bins
        for k, s, p in zip(reversed(list(hp["kernel_sizes"])), reversed(list(hp["strides"])), reversed(list(hp["paddings"]))):
            n = (n - 1) * s - 2 * p + k + (s - 1 if s > 1 else 0)  # + s - 1: the largest length that still maps to n (64 -> 32)
        return n

    def uses_tuned_kernels(self) -> bool:
        """True for the shipped configuration (hidden 17 x 4 levels, 4 bins, k3 s2 p1, 2-16 Gaussians): the tensor-core /
        register-resident kernels of csrc/model.cu + unet_mma.cu.  Every other configuration runs on the generic cell
        kernels (csrc/cell.cu), same results contract, no fused chain."""
        hp = self.hparams
        return (list(hp["hidden_sizes"]) == [17, 17, 17, 17] and hp["num_compressed_bins"] == 4 and set(hp["kernel_sizes"]) == {3}
                and set(hp["strides"]) == {2} and set(hp["paddings"]) == {1} and len(hp["kernel_sizes"]) == 4
                and 2 <= hp["num_gaussians"] <= 16)

    # ---- native model management -------------------------------------------------------------
    def _signature(self):
        # called once per streaming hop: the tensor list is cached (walking the module tree costs ~40 us) and dropped by
        # everything that can replace a Parameter / buffer object (_apply: .to() / .cuda() / .float(); load_state_dict;
        # repack()); storage moves and in-place writes show up in (data_ptr, _version)
        ps = self.__dict__.get("_sig_tensors")
        if ps is None:
            ps = list(self.parameters()) + list(self.buffers())
            self.__dict__["_sig_tensors"] = ps
        return (self._generation,) + tuple([(p.data_ptr(), p._version) for p in ps])

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_sig_tensors"] = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__["_sig_tensors"] = None
        return super().load_state_dict(*args, **kwargs)

    def repack(self) -> None:
        """Force a re-pack on next use.  Needed only after edits autograd cannot see (``p.data.mul_()``, writes through a
        numpy view): ``load_state_dict``, ``optimizer.step`` and ``.to()`` bump the version counters and are picked up
        automatically."""
        with self._native_lock:
            self._generation += 1
            self.__dict__["_sig_tensors"] = None

    def native_model(self, device: torch.device) -> NativeModel:
        """The packed native model for ``device`` (re-packed after any weight change).  Handles are cached per device
        and never destroyed while something still references them: callers that bake device pointers into a CUDA
        graph keep the returned object (``StreamingDenoiser`` does) and compare identities to notice a re-pack."""
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with self._native_lock:
            sig = self._signature()
            cur = self._native.get(idx)
            if cur is not None and cur.signature == sig:
                return cur
            hp = self.hparams
            hs = list(hp["hidden_sizes"])
            uniform = all(h == hs[0] for h in hs)
            ks, ss, pp = set(hp["kernel_sizes"]), set(hp["strides"]), set(hp["paddings"])
            if not (uniform and len(ks) == 1 and len(ss) == 1 and len(pp) == 1):
                raise NotImplementedError(f"GRUUNet2 (B200): non-uniform layer config is not implemented: {hp}")
            cfg = _cabi.ModelConfig(hp["num_compressed_bins"], hs[0], len(hs), ks.pop(), ss.pop(), pp.pop(), hp["num_gaussians"])
            params = [p.detach().to("cpu", torch.float32).contiguous() for p in self.parameters()]
            offs = [g.gs.offset.detach().to("cpu", torch.float32).contiguous()
                    for g in (self.cell.input_gate, self.cell.reset_gate, self.cell.output_gate)]
            parr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            oarr = (C.c_void_p * 3)(*[o.data_ptr() for o in offs])
            handle = C.c_void_p()
            with torch.cuda.device(idx):
                _cabi.check(_cabi.lib().b2d_model_create(C.byref(cfg), parr, len(params), oarr, C.byref(handle)))
            nm = NativeModel(handle, sig, idx)
            self._native[idx] = nm  # the stale one (if any) dies with its last holder
            return nm

    def native_handle(self, device: torch.device):
        return self.native_model(device).handle

    def native_conv_mode(self) -> int:
        """The ``conv_mode`` word of the C-ABI: engine in the low byte, B2D_CONV_EXACT_GATES (0x100) or-ed in."""
        if self.conv_mode not in CONV_MODES:
            raise ValueError(f"conv_mode must be one of {list(CONV_MODES)}, got {self.conv_mode!r}")
        return CONV_MODES[self.conv_mode] | (0x100 if self.exact_gates else 0)

    def _runner(self):
        if self._cell_runner is None:
            from ._cell import ARCH_GRUUNET2, CellRunner

            self._cell_runner = CellRunner(self, ARCH_GRUUNET2)
        return self._cell_runner

    def _wants_grad(self, *tensors) -> bool:
        """Training path: the module is in train() mode, autograd is recording, and something asks for a gradient."""
        return (self.training and torch.is_grad_enabled()
                and (any(t is not None and t.requires_grad for t in tensors) or any(p.requires_grad for p in self.parameters())))

    # ---- forward (gruunet2.py:290-306) -------------------------------------------------------
    def forward(self, input: torch.Tensor, hx: Optional[torch.Tensor] = None):
        """Inference (eval() mode or under ``torch.no_grad()``, as app3.py:200 / server.py:211 call it) runs the fused kernels.
        In train() mode with autograd recording, the call is differentiable: forward and fp32 backward run on the generic cell
        kernels (csrc/cell.cu), gradients reach ``input``, ``hx`` and every parameter (server.py:86-142 ``TrainingContext``)."""
        if self._wants_grad(input, hx):
            return self._forward_train(input, hx)
        with torch.no_grad():
            return self._forward_infer(input, hx)

    def _forward_train(self, input: torch.Tensor, hx: Optional[torch.Tensor]):
        two_dimmed = input.dim() == 2
        if two_dimmed:
            input = input.unsqueeze(0)
        if input.dim() != 3:
            raise Exception(f"unknown!! {input.shape}")
        x = require_cuda_f32(input, "input")
        shape = (x.shape[0], self.latent_size, self.num_compressed_bins)
        h0 = torch.zeros(shape, dtype=x.dtype, device=x.device) if hx is None else require_cuda_f32(hx, "hx")
        if tuple(h0.shape) != shape:
            raise ValueError(f"hx must be {list(shape)}, got {list(h0.shape)}")
        out, h = self._runner().forward_autograd(x, h0, None)
        return (out.squeeze(0) if two_dimmed else out), h

    def _forward_infer(self, input: torch.Tensor, hx: Optional[torch.Tensor] = None):
        two_dimmed = input.dim() == 2
        if two_dimmed:
            input = input.unsqueeze(0)
        if input.dim() != 3:
            raise Exception(f"unknown!! {input.shape}")
        x = require_cuda_f32(input, "input")
        B, T, nm = x.shape
        if self.uses_tuned_kernels() and nm != self.n_mels:
            raise ValueError(f"GRUUNet2 expects {self.n_mels} mel bins on the last axis, got {nm}")
        if hx is None:
            h = torch.zeros(B, self.latent_size, self.num_compressed_bins, dtype=x.dtype, device=x.device)
        else:
            h = require_cuda_f32(hx, "hx").clone()  # the reference never mutates the caller's hx
            if tuple(h.shape) != (B, self.latent_size, self.num_compressed_bins):
                raise ValueError(f"hx must be [{B}, {self.latent_size}, {self.num_compressed_bins}], got {tuple(h.shape)}")
        if not self.uses_tuned_kernels():
            out = self._runner().forward(x, h)
            return (out.squeeze(0) if two_dimmed else out), h
        out = torch.empty_like(x)
        if T == 0:
            return (out.squeeze(0) if two_dimmed else out), h
        mode = self.native_conv_mode()
        lib = _cabi.lib()
        native = self.native_model(x.device)  # held until the launches below are enqueued
        handle = native.handle
        # Re-entrant like the reference module (one model object is shared by every WebRTC session, app3.py:46,449): the
        # packed weights are read-only and the scratch space is per call, taken from torch's stream-ordered caching
        # allocator on the caller's current stream -- two threads / two streams never share it.
        ws = torch.empty(max(int(lib.b2d_gruunet2_workspace_bytes(handle, B, T)), 256), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(lib.b2d_gruunet2_forward(handle, x.data_ptr(), h.data_ptr(), out.data_ptr(), B, T,
                                                 mode, ws.data_ptr(), ws.numel(), stream_ptr(x.device)))
        if two_dimmed:
            out = out.squeeze(0)
        return out, h


class GRUUNet(GRUUNet2):
    """``gruunet.GRUUNet`` (gruunet.py): the same network as GRUUNet2 -- the two files differ only in the class name and an
    unused ``prev`` argument of ``_gruunet`` -- so it shares the kernels."""
