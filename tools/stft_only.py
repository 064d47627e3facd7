"""STFT + mel + log1p only (BASELINE config 2 geometry) for ncu captures and quick timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_denoising_b200 import _cabi, _runtime

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n_fft = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
L = 64000
dev = torch.device("cuda:0")
plan = _runtime.get_plan(n_fft, n_fft // 2, 64, 16000, dev)
T = plan.num_frames(L)
x = torch.randn(B, L, device=dev) * 0.1
peak = torch.full((B,), 0.37, device=dev)
out = torch.empty(B, T, 64, device=dev)
lib = _cabi.lib(); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(scale):
    _cabi.check(lib.b2d_stft_mel_log1p(plan.handle, x.data_ptr(), peak.data_ptr() if scale else None, B, L, out.data_ptr(), None, None, st))
for scale in (True, False):
    run(scale); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(scale); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    alg = B * (4 * L + 4 * 64 * T)
    print(f"B={B} n_fft={n_fft} T={T} scale={scale}: stft+mel+log1p median {t*1e3:.1f} us (L2 flushed), algorithmic {alg/1e6:.1f} MB -> {alg/t/1e6:.0f} GB/s ({alg/t/1e6/6552.3*100:.1f}% of 6552 GB/s)")
