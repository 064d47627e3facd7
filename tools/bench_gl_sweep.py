"""BASELINE config 5: Griffin-Lim-only sweep (n_fft x iterations x batch), achieved algorithmic GB/s per iteration."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_denoising_b200 import _cabi, _runtime

dev = torch.device("cuda:0")
lib = _cabi.lib()
for n_fft in [int(v) for v in os.environ.get('SWEEP_NFFT', '512,1024,2048').split(',')]:
    L = 192000 if n_fft == 1536 else 64000  # 4 s at 48 kHz for the reference-native geometry, 4 s at 16 kHz otherwise
    plan = _runtime.get_plan(n_fft, n_fft // 2, 0, 0, dev)
    T = plan.num_frames(L); F = n_fft // 2 + 1; Lout = plan.out_length(T)
    for B in [int(v) for v in os.environ.get('SWEEP_BATCH', '1,16,256,4096').split(',')]:
        if B * T * plan.frame_stride * 4 * 6 > 60e9:
            continue
        mag = torch.rand(B, T, plan.frame_stride, device=dev)
        wave = torch.empty(B, Lout, device=dev)
        ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(plan.handle, B, T), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        def run(k):
            _cabi.check(lib.b2d_griffinlim_frames(plan.handle, mag.data_ptr(), None, 99, B, T, k, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st))
        def timed(k, reps):
            run(k); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps): run(k)
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        reps = 3 if B >= 256 else 10
        t0 = timed(0, reps)
        for it in [int(v) for v in os.environ.get('SWEEP_ITERS', '32,64,128').split(',')]:
            if B == 4096 and it > 32: continue
            t = timed(it, reps)
            per = (t - t0) / it
            alg = B * (20 * F * T + 8 * Lout)
            print(json.dumps(dict(n_fft=n_fft, batch=B, iters=it, ms_total=round(t, 3), us_per_iter=round(per * 1e3, 2),
                                  L=L, audio_s_per_s=round(B * 4 / (t * 1e-3), 1), gbps=round(alg / per / 1e6, 1),
                                  frac_of_hbm_peak=round(alg / per / 1e6 / 6552.3, 4))), flush=True)
        del mag, wave, ws
        torch.cuda.empty_cache()
