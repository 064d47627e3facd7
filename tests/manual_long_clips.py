"""Run by hand (python tests/manual_long_clips.py): the three register Griffin-Lim paths against the generic kernel on
very long clips (T up to 18751) and on thousands of 4-frame clips; prints SI-SDR, no asserts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_denoising_b200 import _cabi, _runtime
from oracle import metrics
dev = torch.device("cuda:0"); lib = _cabi.lib(); st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(1)
for n_fft in (512, 640, 1024, 1536, 2048):
    plan = _runtime.get_plan(n_fft, n_fft // 2, 0, 0, dev)
    plan_g = _runtime.get_plan(n_fft, n_fft // 2, 0, 0, dev, flags=_runtime.PLAN_GENERIC_KERNELS)
    for B, T in [(2, 3126), (1, 18751), (3000, 4), (3000, 3), (28, 126), (29, 126), (1, 5), (400, 9)]:
        mag = (torch.rand(B, T, plan.frame_stride, generator=g) * 2).to(dev)
        outs = []
        for pl in (plan, plan_g):
            ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(pl.handle, B, T), dtype=torch.uint8, device=dev)
            wave = torch.zeros(B, pl.out_length(T), device=dev)
            _cabi.check(lib.b2d_griffinlim_frames(pl.handle, mag.data_ptr(), None, 0, B, T, 2, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st))
            torch.cuda.synchronize(); outs.append(wave.cpu())
        s = metrics.si_sdr(outs[0], outs[1])
        print(n_fft, B, T, "median", round(float(s.median()), 1), "min", round(float(s.min()), 1), flush=True)
