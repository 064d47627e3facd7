"""The reference-native geometry (app3.py:29-33: 48 kHz, n_fft 1536, hop 768) and the 16 kHz / n_fft 640 geometry of BASELINE
config 3 in batch mode: register-FFT STFT / Griffin-Lim kernels (gl_reg.cu), tensor-core model and inverse mel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audio_denoising_b200 as adb
import bench

dev = torch.device("cuda:0")
sd, cfg = bench.load_model_weights()
m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
for sr, n_fft, B in [(48000, 1536, 64), (48000, 1536, 256), (16000, 640, 256)]:
    L = 4 * sr
    pipe = adb.DenoisePipeline(m, n_fft=n_fft, hop_length=n_fft // 2, n_mels=64, sample_rate=sr, n_iter=32)
    x = bench.synth_batch(B, L, seed=5).to(dev)
    for _ in range(2): pipe.denoise(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): pipe.denoise(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps(dict(sr=sr, n_fft=n_fft, batch=B, clip_s=4, ms_per_batch=round(ms, 3), audio_s_per_s=round(B * 4 / (ms * 1e-3), 1))), flush=True)
