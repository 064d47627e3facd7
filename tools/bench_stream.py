"""BASELINE config 3: streaming real-time mode -- per-hop latency p50/p99 on one B200.

One session (B=1) and a 64-session variant; latency = host-visible time from the chunk being in pinned host memory
to the hop of output samples being back in pinned host memory (H2D, the whole kernel chain with GRU state carried
across hops, D2H, stream sync), i.e. what app3.py:189-215 spends per `while` iteration.
Geometries: 16 kHz n_fft 640 / hop 320 (20 ms hop) and the reference-native 48 kHz n_fft 1536 / hop 768 (16 ms hop).
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import audio_denoising_b200 as adb
from conftest import load_weights

hops = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
warm = 100
dev = torch.device("cuda:0")
sd, cfg = load_weights("dari_tult2")
m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
m.conv_mode = os.environ.get('CONV_MODE', m.conv_mode)
rows = []
for sr, n_fft, hop in [(16000, 640, 320), (48000, 1536, 768), (16000, 1024, 512)]:
    for S in [int(v) for v in os.environ.get('STREAM_SESSIONS', '1,64').split(',')]:
        sdn = adb.StreamingDenoiser(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=sr, sessions=S)
        rng = np.random.default_rng(0)
        sig = (rng.standard_normal((S, n_fft + hop * (hops + warm))) * 0.1).astype(np.float32)
        lat = []
        for i in range(hops + warm):
            win = sig[:, i * hop : i * hop + n_fft]
            t0 = time.perf_counter()
            out = sdn.step(win)
            dt = time.perf_counter() - t0
            if i >= warm:
                lat.append(dt * 1e3)
        lat = np.array(lat)
        row = dict(conv_mode=m.conv_mode, sr=sr, n_fft=n_fft, hop=hop, hop_ms=1000.0 * hop / sr, sessions=S, hops=hops,
                   p50_ms=round(float(np.percentile(lat, 50)), 4), p99_ms=round(float(np.percentile(lat, 99)), 4),
                   mean_ms=round(float(lat.mean()), 4), realtime_factor=round(1000.0 * hop / sr / float(np.percentile(lat, 50)), 1))
        rows.append(row)
        print(json.dumps(row), flush=True)
