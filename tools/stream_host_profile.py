"""Host-side anatomy of StreamingDenoiser.step (graph path): time of each statement, median over many hops."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import audio_denoising_b200 as adb
from audio_denoising_b200.pipeline import draw_seed
from conftest import load_weights

dev = torch.device("cuda:0")
sd, cfg = load_weights("dari_tult2")
m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
n_fft, hop = 640, 320
s = adb.StreamingDenoiser(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=16000, sessions=1)
win = (np.random.default_rng(0).standard_normal((1, n_fft)) * 0.1).astype(np.float32)
for _ in range(50): s.step(win)
names = ["copy_in", "native_model", "seed", "replay", "sync", "copy_out"]
acc = {k: [] for k in names}
pc = time.perf_counter
for _ in range(1000):
    t0 = pc(); s._chunk_host.numpy()[...] = win
    t1 = pc(); ok = s.model.native_model(dev) is s._graph_native
    t2 = pc(); s._seed_host[0] = draw_seed()
    t3 = pc(); s._graph.replay()
    t4 = pc(); torch.cuda.current_stream(dev).synchronize()
    t5 = pc(); out = s._out_host.numpy().copy()
    t6 = pc()
    for k, a, b in zip(names, (t0, t1, t2, t3, t4, t5), (t1, t2, t3, t4, t5, t6)): acc[k].append((b - a) * 1e6)
print({k: round(float(np.median(v)), 1) for k, v in acc.items()}, "us; total", round(sum(float(np.median(v)) for v in acc.values()), 1))
