import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_denoising_b200 import _cabi, _runtime
from oracle import dsp, metrics, synth
dev = torch.device("cuda:0")
n_fft, hop = 2048, 1024
B, L = 5, 20000
x, _ = synth.make_batch(B, L, 16000, start=90)
plan = _runtime.get_plan(n_fft, hop, 64, 16000, dev)
T = plan.num_frames(L)
lib = _cabi.lib(); st = torch.cuda.current_stream().cuda_stream
mag = dsp.stft(x, n_fft, hop).abs()
magd = mag.to(dev)
def gl(seed, n_iter):
    ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(plan.handle, B, T), dtype=torch.uint8, device=dev)
    wave = torch.empty(B, plan.out_length(T), device=dev)
    _cabi.check(lib.b2d_griffinlim(plan.handle, magd.data_ptr(), None, seed, B, T, n_iter, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st))
    return wave.cpu()
for seed, it in [(0, 1), (0, 2), (0, 3), (0, 8), (77, 1), (77, 4), (77, 32)]:
    os.environ.pop("B2D_GL_GENERIC", None)
    f = gl(seed, it)
    os.environ["B2D_GL_GENERIC"] = "1"
    s = gl(seed, it)
    line = f"seed {seed} it {it}: fast vs generic {[round(float(v),1) for v in metrics.si_sdr(f, s)]}"
    if seed == 0:
        ref = dsp.griffinlim(mag, n_fft, hop, it, 0.99, None, rand_init=False)
        line += f" fast vs oracle {[round(float(v),1) for v in metrics.si_sdr(f, ref)]} generic vs oracle {[round(float(v),1) for v in metrics.si_sdr(s, ref)]}"
    print(line)
    d = (f - s).abs()
    print("   per-hop-block max abs diff clip0:", [round(float(d[0, i*hop:(i+1)*hop].max()), 5) for i in range(d.shape[1] // hop)])
