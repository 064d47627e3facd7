"""GRUUNet2 forward only (BASELINE config 2 geometry: 256 clips x 126 frames) per conv_mode: timing + parity vs fp32 mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audio_denoising_b200 as adb
import bench

dev = torch.device("cuda:0")
B, T = 256, 126
sd, cfg = bench.load_model_weights()
m = adb.GRUUNet2(**cfg)
m.load_state_dict(sd)
m = m.to(dev).eval()
x = torch.rand(B, T, 64, device=dev) * 3
ref = None
modes = sys.argv[1:] or ["fp32", "mma", "utc"]
for mode in modes:
    m.conv_mode = mode
    y, h = m(x); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y, h = m(x); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    if ref is None: ref = y.clone()
    rel = float((y - ref).norm() / ref.norm())
    print(f"conv_mode {mode:9s}: forward median {ts[len(ts)//2]*1e3:8.1f} us   rel-L2 vs {modes[0]} {rel:.2e}")
