#!/usr/bin/env python
"""Golden GRADIENTS (fp32 backward), generated FROM THE REFERENCE ITSELF with torch autograd
(build container only: needs /root/reference), like make_golden.py:

  grads.npz : for momo3.MOMO3 (shipped MOMO3-4d4ea0 weights, 24 mel bins), gruunet2.GRUUNet2 (shipped GRUUNet2-good weights)
              and gruunet.GRUUNet (the non-shipped configuration of make_golden_siblings.py, weights from siblings.npz):
              inputs x, h0 and the loss weights wy, wh of   loss = sum(y * wy) + sum(h_T * wh),
              the loss value, and d loss / d x, d loss / d h0, d loss / d every parameter (train() mode; the models have no
              dropout / batch norm, so train() and eval() outputs are identical).
MOMO3 detaches the previous frame of its delta feature (momo3.py:278, 287), so its input gradient is NOT the total derivative
of the forward pass: only autograd through the reference pins it.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_stubs  # noqa: E402


def grads_of(m, x, h0, wy, wh, out, tag):
    m.train()
    for p in m.parameters():
        p.grad = None
    x = x.clone().requires_grad_(True)
    h0 = h0.clone().requires_grad_(True)
    y, h = m(x, h0)
    loss = (y * wy).sum() + (h * wh).sum()
    loss.backward()
    out[f"{tag}_x"], out[f"{tag}_h0"], out[f"{tag}_wy"], out[f"{tag}_wh"] = x.detach().numpy(), h0.detach().numpy(), wy.numpy(), wh.numpy()
    out[f"{tag}_loss"] = np.float64(loss.item())
    out[f"{tag}_gx"], out[f"{tag}_gh0"] = x.grad.numpy(), h0.grad.numpy()
    for k, p in m.named_parameters():
        out[f"{tag}_gp__{k}"] = (p.grad if p.grad is not None else p.detach() * 0).numpy()


def main():
    install_stubs()
    work = tempfile.mkdtemp(prefix="golden_")
    os.symlink(os.path.join(REF, "saves"), os.path.join(work, "saves"))
    os.chdir(work)
    sys.path.insert(0, REF)
    import torch

    torch.set_num_threads(1)
    from gruunet import GRUUNet
    from gruunet2 import GRUUNet2
    from momo3 import MOMO3

    out = {}
    g = torch.Generator().manual_seed(777)
    # MOMO3, shipped weights
    ck = torch.load("saves/MOMO3-4d4ea0/checkpoint.pth", map_location="cpu", weights_only=False)
    m = MOMO3(**ck["config"])
    m.load_state_dict(ck["model_state_dict"])
    x = (torch.randn(2, 7, 24, generator=g).abs() * 1.5).float()
    h0 = (torch.randn(2, 16, 3, generator=g) * 0.5).float()
    grads_of(m, x, h0, torch.randn(2, 7, 24, generator=g), torch.randn(2, 16, 3, generator=g), out, "momo")
    # GRUUNet2, shipped weights
    ck = torch.load("saves/GRUUNet2-good/checkpoint.pth", map_location="cpu", weights_only=False)
    m = GRUUNet2(**ck["config"])
    m.load_state_dict(ck["model_state_dict"])
    x = (torch.rand(3, 6, 64, generator=g) * 2).float()
    h0 = (torch.randn(3, 17, 4, generator=g) * 0.3).float()
    grads_of(m, x, h0, torch.randn(3, 6, 64, generator=g), torch.randn(3, 17, 4, generator=g), out, "g2")
    # GRUUNet, the non-shipped configuration and weights of siblings.npz
    sib = np.load(os.path.join(HERE, "siblings.npz"))
    gcfg = json.loads(bytes(sib["gru_cfg"]).decode())
    gm = GRUUNet(**gcfg)
    gm.load_state_dict({k[len("gru_sd__"):]: torch.from_numpy(sib[k]) for k in sib.files if k.startswith("gru_sd__")})
    x = (torch.randn(2, 5, 20, generator=g).abs() * 1.5).float()
    h0 = (torch.randn(2, 12, 5, generator=g) * 0.5).float()
    grads_of(gm, x, h0, torch.randn(2, 5, 20, generator=g), torch.randn(2, 12, 5, generator=g), out, "gru")
    path = os.path.join(HERE, "grads.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), {t: float(out[f"{t}_loss"]) for t in ("momo", "g2", "gru")})


if __name__ == "__main__":
    main()
