"""Drop-in ``MOMO3`` (reference: momo3.py:246-324; shipped weights ``saves/MOMO3-4d4ea0``) on the generic cell kernels.

GRUUNet2's sibling with a first-order delta feature: every frame enters the cell as two channels ``(x_t, x_t - x_{t-1})``
(momo3.py:277-283), the Gaussian position channels are appended at the encoder input only (momo3.py:129-148), and the decoder
has none (momo3.py:160-190).  Same constructor, ``state_dict`` keys / ``parameters()`` order, ``forward(input, hx=None,
prev=None) -> (out, hx)``, ``hparams`` / ``get_config`` / ``from_config``.  The parameter holders are never called:
``forward`` runs ``b2d_cell_forward`` (csrc/cell.cu).  Inference in eval() / no_grad, differentiable (fp32 backward kernels) in train() mode; CUDA float32 only, no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from ._cell import ARCH_MOMO3, CellRunner
from ._runtime import require_cuda_f32
from .gruunet2 import _PositionCode, _Slot


class _Encoder(nn.Module):  # "downs.{i}.conv", "gs.offset": Gaussian channels on the first layer's input only
    def __init__(self, in_size, sizes, kernels, strides, paddings, count):
        super().__init__()
        chans = [in_size + count] + list(sizes)
        self.downs = nn.ModuleList(
            _Slot(nn.Conv1d(chans[i], chans[i + 1], kernel_size=kernels[i], stride=strides[i], padding=paddings[i])) for i in range(len(sizes)))
        self.gs = _PositionCode(count)


class _Decoder(nn.Module):  # "ups.{i}.conv" (momo3.py:160-179): no Gaussian channels
    def __init__(self, hidden_sizes, kernels, strides, paddings):
        super().__init__()
        rev = list(hidden_sizes)[::-1] + [1]
        self.ups = nn.ModuleList()
        for i in range(len(rev) - 1):
            cin = rev[i] if i == 0 else 2 * rev[i]
            self.ups.append(_Slot(nn.ConvTranspose1d(cin, rev[i + 1], kernel_size=list(kernels)[::-1][i], padding=list(paddings)[::-1][i],
                                                     stride=list(strides)[::-1][i])))


class _Cell(nn.Module):
    def __init__(self, in_size, hidden_sizes, kernel_sizes, strides, paddings, count):
        super().__init__()
        hs = list(hidden_sizes)
        self.input_gate = _Encoder(in_size, hs[:-1] + [3 * hs[-1]], list(kernel_sizes), list(strides), list(paddings), count)
        self.reset_gate = _Encoder(hs[-1], [3 * hs[-1]], [3], [1], [1], count)
        self.output_gate = _Decoder(hs, kernel_sizes, strides, paddings)


class MOMO3(nn.Module):
    def __init__(self, num_compressed_bins, in_size, hidden_sizes, kernel_sizes, strides, paddings, num_gaussians=6):
        super().__init__()
        if in_size != 1:
            raise NotImplementedError("MOMO3 (B200): in_size must be 1 (one mel channel + its delta), as in every shipped configuration")
        self.hparams = dict(num_compressed_bins=num_compressed_bins, in_size=in_size, hidden_sizes=hidden_sizes, kernel_sizes=kernel_sizes,
                            strides=strides, paddings=paddings, num_gaussians=num_gaussians)
        self.latent_size = hidden_sizes[-1]
        self.num_compressed_bins = num_compressed_bins
        self.cell = _Cell(in_size + 1, hidden_sizes, kernel_sizes, strides, paddings, num_gaussians)
        self._generation = 0
        self._runner = CellRunner(self, ARCH_MOMO3)

    def get_config(self):
        return self.hparams

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    def repack(self) -> None:
        self._generation += 1

    def forward(self, input: torch.Tensor, hx: Optional[torch.Tensor] = None, prev: Optional[torch.Tensor] = None):
        """Inference in eval() mode or under ``torch.no_grad()``; differentiable (fp32 backward kernels of csrc/cell.cu) in train()
        mode with autograd recording.  ``prev`` ([B, 1, n_mels] or [B, n_mels]) is the frame before ``input[:, 0]`` and is treated as a
        constant, like the reference's detached copy (momo3.py:278, 287)."""
        train = (self.training and torch.is_grad_enabled()
                 and (input.requires_grad or (hx is not None and hx.requires_grad) or any(p.requires_grad for p in self.parameters())))
        two_dimmed = input.dim() == 2
        if two_dimmed:
            input = input.unsqueeze(0)
        if input.dim() != 3:
            raise Exception(f"unknown!! {input.shape}")
        x = require_cuda_f32(input, "input")
        B = x.shape[0]
        shape = (B, self.latent_size, self.num_compressed_bins)
        if hx is None:
            h = torch.zeros(shape, dtype=x.dtype, device=x.device)
        else:
            h = require_cuda_f32(hx, "hx")
            if tuple(h.shape) != shape:
                raise ValueError(f"hx must be {list(shape)}, got {list(h.shape)}")
        if train:
            out, h = self._runner.forward_autograd(x, h, None if prev is None else prev.detach())
        else:
            with torch.no_grad():
                h = h.clone()  # the reference never mutates the caller's hx
                out = self._runner.forward(x, h, prev)
        return (out.squeeze(0) if two_dimmed else out), h
