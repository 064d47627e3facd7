"""Griffin-Lim only driver (BASELINE config 2 geometry) for ncu captures and quick timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audio_denoising_b200 as adb
from audio_denoising_b200 import _cabi, _runtime

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_fft = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
L = int(sys.argv[4]) if len(sys.argv) > 4 else 64000
n_iter = int(sys.argv[5]) if len(sys.argv) > 5 else 32
dev = torch.device("cuda:0")
plan = _runtime.get_plan(n_fft, n_fft // 2, 0, 0, dev)
T = plan.num_frames(L); F = n_fft // 2 + 1
mag = torch.rand(B, T, plan.frame_stride, device=dev)
init = torch.rand(B, F, T, dtype=torch.complex64, device=dev)
wave = torch.empty(B, plan.out_length(T), device=dev)
lib = _cabi.lib()
ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(plan.handle, B, T), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
def gl(k):
    _cabi.check(lib.b2d_griffinlim_frames(plan.handle, mag.data_ptr(), None, 12345, B, T, k, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st))
def timed(k, reps):
    gl(k); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gl(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
t_full, t_0 = timed(n_iter, reps), timed(0, reps)
per = (t_full - t_0) / max(n_iter, 1)
Lout = plan.out_length(T)
alg = B * (20 * F * T + 8 * Lout)
print(f"B={B} n_fft={n_fft} T={T}: GL {n_iter} it = {t_full:.3f} ms, init+stitch = {t_0:.3f} ms, per iteration {per*1e3:.1f} us, "
      f"algorithmic {alg/1e6:.1f} MB -> {alg/per/1e6:.0f} GB/s ({alg/per/1e6/6552.3*100:.1f}% of 6552 GB/s)")
