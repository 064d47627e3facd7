#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  It imports the unmodified reference modules
``gruunet2`` and ``app3`` under I/O stubs, runs them on seeded inputs on
CPU and writes small ``.npz`` files next to this script.  The tests then compare the oracle
(``oracle/``) and the CUDA path against these files.

Fixtures
  weights_<name>.npz : model_state_dict of saves/GRUUNet2-<name>/checkpoint.pth (+ config as json)
  model_io.npz       : reference GRUUNet2.forward outputs for every checkpoint (full + chunked)
  dsp_chain.npz      : torchaudio transforms built as app3.py:135-153 on a seeded clip, incl.
                       GriffinLim under torch.manual_seed (the init draw is reproduced and saved)
  server_chain.npz   : server.py:207-216 maths with the reference model (noisy-phase path)
  stream.npz         : app3.DenoisingAudioProcessor.recv, hop by hop, seeded per call
"""
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def install_stubs():
    class _Noop:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Noop()

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Noop()

        def __iter__(self):
            return iter(())

        def __bool__(self):
            return False

    for name in ["sounddevice", "matplotlib", "matplotlib.pyplot"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    av = types.ModuleType("av")

    class AudioFrame:
        def __init__(self, arr, sample_rate=48000):
            self._arr = arr
            self.samples = arr.shape[0] if arr.ndim > 1 and arr.shape[1] == 1 else arr.shape[-1]
            self.sample_rate = sample_rate

        def to_ndarray(self, format=None, **kw):
            return self._arr

        @classmethod
        def from_ndarray(cls, arr, format=None, layout=None):
            return cls(arr)

    av.AudioFrame = AudioFrame
    sys.modules["av"] = av

    st = types.ModuleType("streamlit")
    st.cache_resource = lambda f: f

    class _State(dict):
        __getattr__ = dict.get

        def __setattr__(self, k, v):
            self[k] = v

    st.session_state = _State()
    st.sidebar = _Noop()

    def _st_getattr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "columns":
            return lambda spec, *a, **k: [_Noop() for _ in range(spec if isinstance(spec, int) else len(spec))]
        if name == "stop":
            def _stop():
                raise SystemExit
            return _stop
        return _Noop()

    st.__getattr__ = _st_getattr
    sys.modules["streamlit"] = st

    rtc = types.ModuleType("streamlit_webrtc")
    rtc.AudioProcessorBase = object
    rtc.WebRtcMode = _Noop()
    rtc.RTCConfiguration = dict
    rtc.webrtc_streamer = lambda *a, **k: _Noop()
    sys.modules["streamlit_webrtc"] = rtc
    return AudioFrame


def main():
    AudioFrame = install_stubs()
    work = tempfile.mkdtemp(prefix="golden_")
    os.symlink(os.path.join(REF, "saves"), os.path.join(work, "saves"))
    os.chdir(work)  # utils.py:60 creates ./cache in cwd
    sys.path.insert(0, REF)

    import torch
    import torchaudio

    torch.set_num_threads(1)
    from gruunet2 import GRUUNet2  # the reference model, unmodified

    from oracle import synth

    names = ["dari_tult", "dari_tult2", "good"]
    models = {}
    for n in names:
        ck = torch.load(f"saves/GRUUNet2-{n}/checkpoint.pth", map_location="cpu", weights_only=False)
        cfg = {k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in ck["config"].items()}
        sd = ck["model_state_dict"]
        np.savez(
            os.path.join(HERE, f"weights_{n}.npz"),
            __config__=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8),
            **{k: v.numpy() for k, v in sd.items()},
        )
        m = GRUUNet2(**ck["config"])
        m.load_state_dict(sd)
        m.eval()
        models[n] = m

    # ---- model_io: GRUUNet2.forward (gruunet2.py:290-306) -------------------------------------
    g = torch.Generator().manual_seed(42)
    x = (torch.randn(3, 7, 64, generator=g).abs() * 1.5).float()  # log-mel-like, non-negative
    h0 = (torch.randn(3, 17, 4, generator=g) * 0.5).float()
    out = {"x": x.numpy(), "h0": h0.numpy()}
    with torch.no_grad():
        for n, m in models.items():
            y, h = m(x)  # hx=None -> zeros
            out[f"{n}_y"], out[f"{n}_h"] = y.numpy(), h.numpy()
            y2, h2 = m(x, h0)
            out[f"{n}_y_h0"], out[f"{n}_h_h0"] = y2.numpy(), h2.numpy()
            ya, ha = m(x[:, :3])  # chunked carry == full
            yb, hb = m(x[:, 3:], ha)
            assert torch.equal(torch.cat([ya, yb], 1), y) and torch.equal(hb, h)
            y2d, h2d = m(x[0])  # 2-D input path, gruunet2.py:291-293
            out[f"{n}_y2d"], out[f"{n}_h2d"] = y2d.numpy(), h2d.numpy()
    np.savez(os.path.join(HERE, "model_io.npz"), **out)

    # ---- dsp_chain: the transforms exactly as app3.py:135-153 builds them -------------------
    chain = {}
    for tag, (n_fft, hop, sr, L) in {"a": (512, 256, 16000, 4096), "b": (1024, 512, 16000, 6000), "c": (640, 320, 16000, 2560)}.items():
        noisy, _ = synth.make_batch(2, L, sr, start=100)
        T0 = torchaudio.transforms.Spectrogram(power=None, n_fft=n_fft, win_length=n_fft, hop_length=hop, window_fn=torch.hann_window)
        M0T = torchaudio.transforms.MelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr)
        M0I = torchaudio.transforms.InverseMelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr)
        GL = torchaudio.transforms.GriffinLim(n_fft=n_fft, win_length=n_fft, hop_length=hop, window_fn=torch.hann_window, power=1.0)
        I0 = torchaudio.transforms.InverseSpectrogram(n_fft=n_fft, win_length=n_fft, hop_length=hop)
        m = models["good"]
        with torch.no_grad():
            spec = T0(noisy)
            logmel = M0T(spec.abs()).log1p()
            feats = logmel.transpose(-1, -2)
            pred, hx = m(feats, None)
            rec = torch.nn.functional.leaky_relu(feats - pred, negative_slope=0.2)
            mel_mag = torch.clamp(torch.expm1(rec.transpose(-1, -2)), min=0)
            lin = torch.clamp(M0I(mel_mag), min=0)
            torch.manual_seed(7)
            wave = GL(lin)
            torch.manual_seed(7)
            init = torch.rand(lin.size(), dtype=torch.complex64)  # the draw at TA functional.py:310
            rt = I0(spec)
        chain.update(
            {
                f"{tag}_cfg": np.array([n_fft, hop, sr, L]),
                f"{tag}_noisy": noisy.numpy(),
                f"{tag}_spec": spec.numpy(),
                f"{tag}_logmel": logmel.numpy(),
                f"{tag}_pred": pred.numpy(),
                f"{tag}_lin": lin.numpy(),
                f"{tag}_init": init.numpy(),
                f"{tag}_wave": wave.numpy(),
                f"{tag}_istft": rt.numpy(),
            }
        )
    np.savez_compressed(os.path.join(HERE, "dsp_chain.npz"), **chain)

    # ---- server_chain: server.py:207-216 with the reference model ("GRUUNet2-good", SR=48000) ---
    n_fft, hop, sr = 1024, 512, 48000
    x, _ = synth.make_batch(1, 6144, sr, start=200)
    T0 = torchaudio.transforms.Spectrogram(power=None, n_fft=n_fft, win_length=n_fft, hop_length=hop)
    I0 = torchaudio.transforms.InverseSpectrogram(n_fft=n_fft, win_length=n_fft, hop_length=hop)
    M0T = torchaudio.transforms.MelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr)
    M0I = torchaudio.transforms.InverseMelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr)
    m = models["good"]
    hx = None
    srv = {"x": x.numpy()}
    for rep in range(2):  # two requests: hx carried with the 0.9 leak
        abs_spec = T0(x)
        phase = abs_spec.angle()
        magn = abs_spec.abs()
        log_mel_mag = M0T(magn).log1p()
        with torch.no_grad():
            o, hx = m(log_mel_mag.transpose(-1, -2), hx)
            o = torch.nn.functional.leaky_relu(o.transpose(-1, -2), negative_slope=0) * 3
            hx = hx * 0.9
        O = M0I((log_mel_mag - o).exp() - 1)
        W = I0(torch.polar(O, phase))
        srv[f"wave{rep}"] = W.numpy()
        srv[f"hx{rep}"] = hx.numpy()
    np.savez_compressed(os.path.join(HERE, "server_chain.npz"), **srv)

    # ---- stream: app3.DenoisingAudioProcessor.recv under stubs --------------------------------
    import app3  # noqa: E402  (loads saves/GRUUNet2-dari_tult2 via its own loader)

    assert app3.model is not None, "reference loader failed"
    stream = {}
    for tag, (n_fft, hop, sr, nhops) in {"s16k": (640, 320, 16000, 8), "s48k": (1536, 768, 48000, 5)}.items():
        proc = app3.DenoisingAudioProcessor(app3.model, app3.device, app3.GRUUNET2_CONFIG, {"n_fft": n_fft, "hop_length": hop, "n_mels": 64}, sr)
        sig, _ = synth.make_batch(1, hop * (nhops + 2), sr, start=300)
        pcm = (sig[0].numpy() * 0.8 * 32767).astype(np.int16)
        outs = []
        for i in range(nhops + 2):
            torch.manual_seed(1000 + i)
            fr = proc.recv(AudioFrame(pcm[i * hop : (i + 1) * hop].reshape(-1, 1), sr))
            outs.append(np.asarray(fr._arr).reshape(-1))
        stream[f"{tag}_cfg"] = np.array([n_fft, hop, sr, nhops + 2])
        stream[f"{tag}_pcm"] = pcm
        stream[f"{tag}_out"] = np.stack(outs)
        stream[f"{tag}_hx"] = proc.hx.numpy()
    np.savez_compressed(os.path.join(HERE, "stream.npz"), **stream)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
