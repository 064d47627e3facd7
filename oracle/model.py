"""Oracle GRUUNet2 (torch CPU, fp32) -- test infrastructure, see oracle/__init__.py.

Functional restatement of gruunet2.py:54-306 driven directly by a reference
``model_state_dict`` (keys listed in SURVEY.md §8b).  It keeps the reference's
*computational structure* -- one cell evaluation per time frame, Gaussian position
channels rebuilt and concatenated before every convolution -- so that timing it on
the host CPU (bench.py ``cpu_baseline``) is representative of the reference.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch
import torch.nn.functional as F

_IN = "cell.input_gate"
_RS = "cell.reset_gate"
_OUT = "cell.output_gate"


def default_config() -> dict:
    """GRUUNET2_CONFIG of app3.py:18-26 == ``config`` of the three shipped checkpoints."""
    return dict(
        num_compressed_bins=4,
        in_size=1,
        hidden_sizes=(17, 17, 17, 17),
        kernel_sizes=(3, 3, 3, 3),
        strides=(2, 2, 2, 2),
        paddings=(1, 1, 1, 1),
        num_gaussians=6,
    )


def random_state_dict(config: Mapping | None = None, seed: int = 0) -> dict:
    """Seeded random weights with the reference's tensor names/shapes.

    Uses the same fan-in uniform bound as nn.Conv1d's default init so magnitudes are
    realistic; it is NOT meant to reproduce torch's RNG stream.
    """
    cfg = dict(default_config() if config is None else config)
    g = torch.Generator().manual_seed(seed)
    hs = list(cfg["hidden_sizes"])
    G = cfg["num_gaussians"]
    ks = list(cfg["kernel_sizes"])
    sd: dict[str, torch.Tensor] = {}

    def uni(shape, fan_in):
        b = 1.0 / fan_in**0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    enc_out = hs[:-1] + [3 * hs[-1]]
    enc_in = [cfg["in_size"]] + hs[:-1]
    for i, (ci, co) in enumerate(zip(enc_in, enc_out)):
        sd[f"{_IN}.downs.{i}.conv.weight"] = uni((co, ci + G, ks[i]), (ci + G) * ks[i])
        sd[f"{_IN}.downs.{i}.conv.bias"] = uni((co,), (ci + G) * ks[i])
    sd[f"{_IN}.gs.offset"] = torch.linspace(0.0, 1.0, G)
    sd[f"{_RS}.downs.0.conv.weight"] = uni((3 * hs[-1], hs[-1] + G, 3), (hs[-1] + G) * 3)
    sd[f"{_RS}.downs.0.conv.bias"] = uni((3 * hs[-1],), (hs[-1] + G) * 3)
    sd[f"{_RS}.gs.offset"] = torch.linspace(0.0, 1.0, G)
    sizes = [1] + hs  # UpBlocks: sizes = [output_size, *hidden_sizes], walked backwards
    rs = sizes[::-1]
    rk = ks[::-1]
    for i in range(len(hs)):
        ci = rs[i] + G if i == 0 else 2 * rs[i] + G
        co = rs[i + 1]
        sd[f"{_OUT}.ups.{i}.conv.weight"] = uni((ci, co, rk[i]), co * rk[i])
        sd[f"{_OUT}.ups.{i}.conv.bias"] = uni((co,), co * rk[i])
    sd[f"{_OUT}.gs.offset"] = torch.linspace(0.0, 1.0, G)
    return sd


class GRUUNet2Oracle:
    """forward(x[B,T,n_mels] or [T,n_mels], hx=None) -> (out, hx[B,H,bins]) -- gruunet2.py:290-306."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], config: Mapping | None = None):
        cfg = dict(default_config() if config is None else config)
        assert cfg["in_size"] == 1  # gruunet2.py:257
        self.cfg = cfg
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
        self.levels = len(cfg["hidden_sizes"])
        self.hidden = int(cfg["hidden_sizes"][-1])
        self.bins = int(cfg["num_compressed_bins"])
        self.strides: Sequence[int] = list(cfg["strides"])
        self.paddings: Sequence[int] = list(cfg["paddings"])

    # gruunet2.py:54-68 -- exp(coeff * (p - o)^2), p = linspace(0,1,n), coeff from the f32 offset gap
    def _position_channels(self, prefix: str, n: int, batch: int) -> torch.Tensor:
        off = self.sd[f"{prefix}.gs.offset"]
        coeff = -0.5 / (off[1] - off[0]).item() ** 2
        pos = torch.linspace(0, 1, n)
        g = torch.exp(coeff * (pos[:, None] - off[None, :]) ** 2)  # [n, G]
        return g.t().unsqueeze(0).expand(batch, -1, -1)  # [B, G, n]

    # gruunet2.py:127-144 (return_samples=True): list [x, d0, d1, d2, gates_x]
    def _encode(self, x: torch.Tensor) -> list[torch.Tensor]:
        feats = [x]
        for i in range(self.levels):
            cur = feats[-1]
            inp = torch.cat([cur, self._position_channels(_IN, cur.shape[-1], cur.shape[0])], dim=1)
            y = F.conv1d(
                inp,
                self.sd[f"{_IN}.downs.{i}.conv.weight"],
                self.sd[f"{_IN}.downs.{i}.conv.bias"],
                stride=self.strides[i],
                padding=self.paddings[i],
            )
            feats.append(torch.relu(y))
        return feats

    # gruunet2.py:146-155 + 218-222: one stride-1 conv on the hidden state, ReLU'd
    def _hidden_gates(self, h: torch.Tensor) -> torch.Tensor:
        inp = torch.cat([h, self._position_channels(_RS, h.shape[-1], h.shape[0])], dim=1)
        y = F.conv1d(inp, self.sd[f"{_RS}.downs.0.conv.weight"], self.sd[f"{_RS}.downs.0.conv.bias"], stride=1, padding=1)
        return torch.relu(y)

    # gruunet2.py:184-199 + 81-96: transposed convs, relu + skip concat except on the last
    def _decode(self, feats: list[torch.Tensor]) -> torch.Tensor:
        h = feats[-1]
        rs = list(self.strides)[::-1]
        rp = list(self.paddings)[::-1]
        for i in range(self.levels):
            skip = feats[self.levels - 1 - i]
            inp = torch.cat([h, self._position_channels(_OUT, h.shape[-1], h.shape[0])], dim=1)
            w = self.sd[f"{_OUT}.ups.{i}.conv.weight"]
            k = w.shape[-1]
            want = skip.shape[-1]
            base = (inp.shape[-1] - 1) * rs[i] - 2 * rp[i] + k
            y = F.conv_transpose1d(
                inp, w, self.sd[f"{_OUT}.ups.{i}.conv.bias"], stride=rs[i], padding=rp[i], output_padding=want - base
            )
            h = y if i == self.levels - 1 else torch.cat([torch.relu(y), skip], dim=1)
        return h

    # gruunet2.py:228-244
    def cell(self, x_t: torch.Tensor, h: torch.Tensor):
        feats = self._encode(x_t)
        gx = feats[-1]
        gh = self._hidden_gates(h)
        H = self.hidden
        xr, xz, xn = gx[:, :H], gx[:, H : 2 * H], gx[:, 2 * H :]
        hr, hz, hn = gh[:, :H], gh[:, H : 2 * H], gh[:, 2 * H :]
        z = torch.sigmoid(xz + hz)
        r = torch.sigmoid(xr + hr)
        n = torch.tanh(xn + r * hn)
        h_new = n + z * (h - n)
        out = self._decode(feats[:-1] + [h_new]).squeeze(-2)
        return out, h_new

    @torch.no_grad()
    def forward(self, x: torch.Tensor, hx: torch.Tensor | None = None):
        squeeze = x.dim() == 2
        if squeeze:
            x = x.unsqueeze(0)
        if hx is None:
            hx = torch.zeros(x.shape[0], self.hidden, self.bins, dtype=x.dtype)
        outs = []
        for t in range(x.shape[1]):
            o, hx = self.cell(x[:, t, :].unsqueeze(1), hx)
            outs.append(o)
        out = torch.stack(outs, dim=1)
        return (out.squeeze(0) if squeeze else out), hx

    __call__ = forward
