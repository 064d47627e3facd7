"""Host side of ``b2d_cell`` (csrc/cell.cu): the reference's conv-GRU U-Net cell for any configuration.

``CellRunner`` packs a module's parameters into a native cell (one per device, input length and weight version; immutable,
kept alive by whoever holds it) and runs ``b2d_cell_forward``.  Used by ``MOMO3`` and by ``GRUUNet2`` / ``GRUUNet`` whenever the
configuration is not the shipped one the tuned kernels are written for.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref

import torch

from . import _cabi
from ._runtime import ptr, require_cuda_f32, stream_ptr

ARCH_GRUUNET2, ARCH_MOMO3 = 0, 1


class NativeCell:
    def __init__(self, handle, key):
        self.handle, self.key = handle, key
        self._fin = weakref.finalize(self, _cabi.lib().b2d_cell_destroy, handle)


class CellRunner:
    def __init__(self, module: torch.nn.Module, arch: int):
        self._module = weakref.ref(module)
        self.arch = arch
        self._cells: dict = {}
        self._lock = threading.Lock()

    def _signature(self, module):
        ps = list(module.parameters()) + list(module.buffers())
        return (getattr(module, "_generation", 0),) + tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)

    def native(self, device: torch.device, n_mels: int) -> NativeCell:
        module = self._module()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with self._lock:
            key = (idx, n_mels, self._signature(module))
            cur = self._cells.get((idx, n_mels))
            if cur is not None and cur.key == key:
                return cur
            hp = module.hparams
            hs, ks, ss, pp = (list(hp[k]) for k in ("hidden_sizes", "kernel_sizes", "strides", "paddings"))
            if not (len(hs) == len(ks) == len(ss) == len(pp)) or not 1 <= len(hs) <= 8:
                raise ValueError(f"hidden_sizes / kernel_sizes / strides / paddings must have the same length in [1, 8]: {hp}")
            cfg = _cabi.CellConfig()
            cfg.arch, cfg.num_compressed_bins, cfg.levels = self.arch, int(hp["num_compressed_bins"]), len(hs)
            cfg.num_gaussians, cfg.n_mels = int(hp["num_gaussians"]), int(n_mels)
            for i in range(len(hs)):
                cfg.hidden[i], cfg.kernel[i], cfg.stride[i], cfg.padding[i] = int(hs[i]), int(ks[i]), int(ss[i]), int(pp[i])
            params = [p.detach().to("cpu", torch.float32).contiguous() for p in module.parameters()]
            gates = [module.cell.input_gate, module.cell.reset_gate] + ([module.cell.output_gate] if self.arch == ARCH_GRUUNET2 else [])
            offs = [g.gs.offset.detach().to("cpu", torch.float32).contiguous() for g in gates]
            parr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            oarr = (C.c_void_p * 3)(*([o.data_ptr() for o in offs] + [None] * (3 - len(offs))))
            handle = C.c_void_p()
            with torch.cuda.device(idx):
                _cabi.check(_cabi.lib().b2d_cell_create(C.byref(cfg), parr, len(params), oarr, C.byref(handle)))
            nc = NativeCell(handle, key)
            self._cells[(idx, n_mels)] = nc
            return nc

    def forward_autograd(self, x: torch.Tensor, h0: torch.Tensor, prev: torch.Tensor | None = None):
        """Differentiable forward: (out [B, T, n_mels], h_T [B, H, bins]) with gradients to x, h0 and every parameter of the
        module, computed by the fp32 backward kernels of csrc/cell.cu (``b2d_cell_backward``)."""
        module = self._module()
        params = [p for p in module.parameters()]
        return _CellFunction.apply(self, x, h0, prev, *params)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, h: torch.Tensor, prev: torch.Tensor | None = None):
        """x [B, T, n_mels], h [B, H, bins] (updated in place), prev [B, n_mels] or None -> out [B, T, n_mels]."""
        B, T, nm = x.shape
        out = torch.empty_like(x)
        if T == 0:
            return out
        lib = _cabi.lib()
        cell = self.native(x.device, nm)
        if prev is not None:
            prev = require_cuda_f32(prev, "prev").reshape(B, nm)
        ws = torch.empty(max(int(lib.b2d_cell_workspace_bytes(cell.handle, B, T)), 256), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(lib.b2d_cell_forward(cell.handle, x.data_ptr(), ptr(prev), h.data_ptr(), out.data_ptr(), B, T,
                                             ws.data_ptr(), ws.numel(), stream_ptr(x.device)))
        return out


class _CellFunction(torch.autograd.Function):
    """autograd bridge: forward = b2d_cell_forward (its workspace is kept for the backward), backward = b2d_cell_backward."""

    @staticmethod
    def forward(ctx, runner: CellRunner, x, h0, prev, *params):
        x = require_cuda_f32(x, "input")
        h0 = require_cuda_f32(h0, "hx")
        B, T, nm = x.shape
        lib = _cabi.lib()
        cell = runner.native(x.device, nm)
        if prev is not None:
            prev = require_cuda_f32(prev, "prev").reshape(B, nm)
        h = h0.clone()
        out = torch.empty_like(x)
        ws = torch.empty(max(int(lib.b2d_cell_workspace_bytes(cell.handle, B, max(T, 1))), 256), dtype=torch.uint8, device=x.device)
        if T > 0:
            with torch.cuda.device(x.device):
                _cabi.check(lib.b2d_cell_forward(cell.handle, x.data_ptr(), ptr(prev), h.data_ptr(), out.data_ptr(), B, T,
                                                 ws.data_ptr(), ws.numel(), stream_ptr(x.device)))
        ctx.cell, ctx.ws, ctx.prev = cell, ws, prev
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(x, h0)
        ctx.mark_non_differentiable()
        return out, h

    @staticmethod
    def backward(ctx, grad_out, grad_h):
        x, h0 = ctx.saved_tensors
        B, T, nm = x.shape
        lib = _cabi.lib()
        cell = ctx.cell
        n = lib.b2d_cell_num_param_floats(cell.handle)
        gx = torch.zeros_like(x)
        gh0 = torch.zeros_like(h0) if grad_h is None else grad_h.contiguous().clone()
        gp = torch.zeros(n, dtype=torch.float32, device=x.device)
        if T > 0:
            go = (torch.zeros_like(x) if grad_out is None else grad_out.contiguous().float())
            gh = None if grad_h is None else grad_h.contiguous().float()
            ws = torch.empty(max(int(lib.b2d_cell_backward_workspace_bytes(cell.handle, B, T)), 256), dtype=torch.uint8, device=x.device)
            with torch.cuda.device(x.device):
                _cabi.check(lib.b2d_cell_backward(cell.handle, x.data_ptr(), ptr(ctx.prev), h0.data_ptr(), ctx.ws.data_ptr(), go.data_ptr(),
                                                  ptr(gh), gx.data_ptr(), gh0.data_ptr(), gp.data_ptr(), B, T, ws.data_ptr(), ws.numel(),
                                                  stream_ptr(x.device)))
        grads, o = [], 0
        for shp in ctx.shapes:
            k = 1
            for v in shp:
                k *= v
            grads.append(gp[o:o + k].view(shp))
            o += k
        return (None, gx, gh0, None, *grads)
