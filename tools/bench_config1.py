"""BASELINE config 1: ONE 4 s / 16 kHz noisy clip through STFT -> Mel -> GRUUNet2 -> inverse Mel -> Griffin-Lim (32 it) -> iSTFT.

CPU leg: the oracle port of the reference path (torchaudio arithmetic + the reference's per-frame model loop), best of 5
after one warm-up, at torch.set_num_threads(1) and at os.cpu_count().  GPU leg: the same clip through
DenoisePipeline.denoise (B = 1), device-resident and host-to-host, median of 200.
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audio_denoising_b200 as adb
import bench  # the CPU leg is bench.py's own cpu_baseline code path (bench.cpu_pass_seconds): the oracle port, timed on the host cores

L = 64000
noisy = bench.synth_batch(1, L, seed=1234)
sd, cfg = bench.load_model_weights()
row = {"config": "1 clip x 4 s @ 16 kHz, n_fft 1024, hop 512, 64 mels, GL 32 it (BASELINE configs[0])", "cpu_count": os.cpu_count()}
arm = bench.CpuArm()
pcm1 = bench.to_pcm16(noisy)
for threads in (1, os.cpu_count()):
    best = arm.seconds(pcm1, 5, threads)
    row[f"cpu_{threads}_threads_ms"] = round(best * 1e3, 2)
    row[f"cpu_{threads}_threads_audio_s_per_s"] = round(4.0 / best, 1)
dev = torch.device("cuda:0")
m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=32)
xd = noisy.to(dev)
for _ in range(10): pipe.denoise(xd)
torch.cuda.synchronize()
ts = []
for _ in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pipe.denoise(xd); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
row["gpu_device_resident_ms_p50"] = round(ts[100], 4)
hp = noisy.pin_memory()
out = torch.empty((1, pipe.out_length(L)), dtype=torch.float32, pin_memory=True)
for _ in range(10): pipe.denoise_host(hp, out)
hs = []
for _ in range(200):
    t0 = time.perf_counter(); pipe.denoise_host(hp, out); hs.append((time.perf_counter() - t0) * 1e3)
hs.sort()
row["gpu_host_to_host_ms_p50"] = round(hs[100], 4)
row["gpu_audio_s_per_s_single_clip"] = round(4.0 / (hs[100] * 1e-3), 1)
print(json.dumps(row), flush=True)
