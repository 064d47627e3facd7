"""First-contact GPU check (run by hand: python tests/manual_gpu_check.py): every stage against the oracle, errors printed, no asserts."""
import os, sys, time, json, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import audio_denoising_b200 as adb
from audio_denoising_b200 import _cabi, _runtime
from oracle import dsp, metrics, model as omodel, pipeline as opipe, synth
from conftest import load_weights

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), "lib version", adb.native_library().b2d_version())

def stage(name):
    def deco(fn):
        try:
            t = time.time(); fn(); torch.cuda.synchronize(); print(f"[ok ] {name} ({time.time()-t:.2f}s)")
        except Exception as e:
            print(f"[ERR] {name}: {e}"); traceback.print_exc()
        return fn
    return deco

@stage("stft")
def _():
    for n_fft, hop in [(1024, 512), (512, 256), (640, 320), (1536, 768), (2048, 1024), (400, 100)]:
        x, _ = synth.make_batch(2, 5000, 16000)
        ref = dsp.stft(x, n_fft, hop)
        got = adb.Spectrogram(power=None, n_fft=n_fft, hop_length=hop).to(dev)(x.to(dev)).cpu()
        print("   stft", n_fft, hop, "rel", metrics.rel_l2(got, ref))

@stage("logmel")
def _():
    for n_fft, hop, sr in [(1024, 512, 16000), (1536, 768, 48000)]:
        x, _ = synth.make_batch(2, 6000, sr)
        fb = dsp.mel_fbanks(n_fft // 2 + 1, 64, sr)
        ref = dsp.log_mel(x, n_fft, hop, fb)
        plan = _runtime.get_plan(n_fft, hop, 64, sr, dev)
        T = plan.num_frames(6000)
        bm = torch.empty(2, 64, T, device=dev)
        _cabi.check(_cabi.lib().b2d_stft_mel_log1p(plan.handle, x.to(dev).data_ptr(), None, 2, 6000, None, bm.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
        print("   logmel", n_fft, "rel", metrics.rel_l2(bm.cpu(), ref), "fb equal", torch.equal(plan.fb, fb))

@stage("inverse mel")
def _():
    mel = torch.rand(2, 64, 37) * 4
    fb = dsp.mel_fbanks(513, 64, 16000)
    ref = dsp.inverse_mel(mel, fb)
    got = adb.InverseMelScale(n_mels=64, n_stft=513, sample_rate=16000).to(dev)(mel.to(dev)).cpu()
    print("   invmel rel", metrics.rel_l2(got, ref))

@stage("istft")
def _():
    for n_fft, hop in [(1024, 512), (640, 320)]:
        x, _ = synth.make_batch(2, hop * 13, 16000)
        spec = dsp.stft(x, n_fft, hop)
        got = adb.InverseSpectrogram(n_fft=n_fft, hop_length=hop).to(dev)(spec.to(dev)).cpu()
        print("   istft", n_fft, "rel vs x", metrics.rel_l2(got, x))

@stage("model")
def _():
    z = np.load(os.path.join(ROOT, "tests/golden/model_io.npz"))
    for name in ["dari_tult", "good"]:
        sd, cfg = load_weights(name)
        m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev)
        y, h = m(torch.from_numpy(z["x"]).to(dev))
        print("   model", name, "y rel", metrics.rel_l2(y.cpu(), torch.from_numpy(z[name + "_y"])), "h rel", metrics.rel_l2(h.cpu(), torch.from_numpy(z[name + "_h"])))

@stage("griffin-lim")
def _():
    for n_fft, hop, L, B in [(1024, 512, 16000, 2), (640, 320, 6400, 2), (1536, 768, 1536, 2), (1024, 512, 64000, 2)]:
        x, _ = synth.make_batch(B, L, 16000)
        mag = dsp.stft(x, n_fft, hop).abs()
        init = synth.gl_init_angles(mag.shape)
        for n_iter in (0, 1, 32):
            ref = dsp.griffinlim(mag, n_fft, hop, n_iter, 0.99, init)
            got = adb.GriffinLim(n_fft=n_fft, hop_length=hop, power=1.0, n_iter=n_iter).to(dev)(mag.to(dev), init_angles=init.to(dev)).cpu()
            print("   GL", n_fft, L, "it", n_iter, "SI-SDR", [round(float(v), 1) for v in metrics.si_sdr(got, ref)])

@stage("pipeline")
def _():
    sd, cfg = load_weights("good")
    m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev)
    noisy, clean = synth.make_batch(2, 16000, 16000)
    init = synth.gl_init_angles((2, 513, 32))
    ref = opipe.denoise_batch(noisy, omodel.GRUUNet2Oracle(sd, cfg), 1024, 512, 64, 16000, 32, 0.99, init)
    pipe = adb.DenoisePipeline(m, 1024, 512, 64, 16000)
    r = pipe.denoise(noisy.to(dev), init_angles=init.to(dev), return_intermediates=True)
    print("   logmel", metrics.rel_l2(r["logmel"].cpu(), ref["logmel"]), "pred", metrics.rel_l2(r["pred"].cpu(), ref["pred"]),
          "lin", metrics.rel_l2(r["lin_mag"].cpu(), ref["lin_mag"]), "wave SI-SDR", metrics.si_sdr(r["wave"].cpu(), ref["wave"]).tolist())

@stage("timing B=256")
def _():
    sd, cfg = load_weights("good")
    m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev)
    pipe = adb.DenoisePipeline(m, 1024, 512, 64, 16000)
    x = torch.rand(256, 64000, device=dev) * 2 - 1
    init = torch.rand(256, 513, 126, dtype=torch.complex64, device=dev)
    for _ in range(2): pipe.denoise(x, init_angles=init)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): pipe.denoise(x, init_angles=init)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"   256 x 4 s: {ms:.2f} ms/step -> {256*4/ms*1e3:.0f} audio-s/s")
