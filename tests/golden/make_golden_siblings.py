#!/usr/bin/env python
"""Golden fixtures for the sibling models, generated FROM THE REFERENCE ITSELF like
make_golden.py (build container only: needs /root/reference):

  weights_momo3.npz : model_state_dict + config of saves/MOMO3-4d4ea0/checkpoint.pth (momo3.py:246-324)
  siblings.npz      : momo3.MOMO3.forward with those weights on inputs of 24 and 27 mel bins (both compress to the model's 3
                      bins: paddings (1, 0, 1), odd lengths), full and chunked with hx / prev carried;
                      gruunet.GRUUNet.forward (gruunet.py: the same cell as gruunet2) for a NON-shipped configuration
                      (hidden (8, 12), kernels (3, 5), strides (2, 2), paddings (1, 2), 5 bins, 4 Gaussians), seeded weights
                      saved beside the outputs.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_stubs  # noqa: E402


def main():
    install_stubs()
    work = tempfile.mkdtemp(prefix="golden_")
    os.symlink(os.path.join(REF, "saves"), os.path.join(work, "saves"))
    os.chdir(work)
    sys.path.insert(0, REF)
    import torch

    torch.set_num_threads(1)
    from gruunet import GRUUNet
    from momo3 import MOMO3

    out = {}
    ck = torch.load("saves/MOMO3-4d4ea0/checkpoint.pth", map_location="cpu", weights_only=False)
    cfg = {k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in ck["config"].items()}
    np.savez(os.path.join(HERE, "weights_momo3.npz"), __config__=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8),
             **{k: v.numpy() for k, v in ck["model_state_dict"].items()})
    m = MOMO3(**ck["config"])
    m.load_state_dict(ck["model_state_dict"])
    m.eval()
    g = torch.Generator().manual_seed(4242)
    with torch.no_grad():
        for nm in (24, 27):
            x = (torch.randn(2, 9, nm, generator=g).abs() * 1.5).float()
            h0 = (torch.randn(2, 16, 3, generator=g) * 0.5).float()
            y, h = m(x)
            y2, h2 = m(x, h0)
            ya, ha = m(x[:, :4])
            yb, hb = m(x[:, 4:], ha, prev=x[:, 3:4].transpose(1, 2).transpose(1, 2))  # prev: the frame before the chunk, [B, 1, n_mels]
            assert torch.allclose(torch.cat([ya, yb], 1), y, atol=1e-6) and torch.allclose(hb, h, atol=1e-6)
            y2d, h2d = m(x[0])
            out.update({f"momo_{nm}_x": x.numpy(), f"momo_{nm}_h0": h0.numpy(), f"momo_{nm}_y": y.numpy(), f"momo_{nm}_h": h.numpy(),
                        f"momo_{nm}_y_h0": y2.numpy(), f"momo_{nm}_h_h0": h2.numpy(), f"momo_{nm}_y2d": y2d.numpy(), f"momo_{nm}_h2d": h2d.numpy()})
    gcfg = dict(num_compressed_bins=5, in_size=1, hidden_sizes=[8, 12], kernel_sizes=[3, 5], strides=[2, 2], paddings=[1, 2], num_gaussians=4)
    torch.manual_seed(99)
    gm = GRUUNet(**gcfg)
    gm.eval()
    x = (torch.randn(3, 6, 20, generator=g).abs() * 1.5).float()
    with torch.no_grad():
        y, h = gm(x)
    out["gru_cfg"] = np.frombuffer(json.dumps(gcfg).encode(), dtype=np.uint8)
    out["gru_x"], out["gru_y"], out["gru_h"] = x.numpy(), y.numpy(), h.numpy()
    for k, v in gm.state_dict().items():
        out["gru_sd__" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "siblings.npz"), **out)
    for f in ("weights_momo3.npz", "siblings.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
