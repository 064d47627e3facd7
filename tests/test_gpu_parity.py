"""GPU parity tests: the CUDA path (through the Python host layer -> C-ABI) against the CPU oracle.

Run on the B200 box with ``pytest -m gpu``.  Tolerances (north_star): STFT / Mel within 1e-4 relative
(fp32); model fp32 mode rel-L2 <= 1e-5; Griffin-Lim with identical injected initial angles
SI-SDR(ours, oracle) >= 60 dB and |SI-SDR vs clean difference| <= 0.05 dB.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import CHECKPOINTS, load_weights

pytestmark = pytest.mark.gpu

GEOMS = [(1024, 512), (512, 256), (2048, 1024), (640, 320), (1536, 768), (256, 128)]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import audio_denoising_b200 as adb

    adb.native_library()  # must load: no fallback
    return torch.device("cuda:0")


def _oracle():
    from oracle import dsp, metrics, model, pipeline, synth

    return dsp, metrics, model, pipeline, synth


def _our_model(name, dev):
    import audio_denoising_b200 as adb

    sd, cfg = load_weights(name)
    m = adb.GRUUNet2(**cfg)
    m.load_state_dict(sd)
    return m.to(dev).eval(), sd, cfg


# ---------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n_fft,hop", GEOMS)
@pytest.mark.parametrize("L", [4096, 5000])
def test_stft_matches_oracle(dev, n_fft, hop, L):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    x, _ = synth.make_batch(3, L, 16000, start=10)
    ref = dsp.stft(x, n_fft, hop)
    T0 = adb.Spectrogram(power=None, n_fft=n_fft, win_length=n_fft, hop_length=hop, window_fn=torch.hann_window).to(dev)
    got = T0(x.to(dev)).cpu()
    assert got.shape == ref.shape and got.dtype == torch.complex64
    assert metrics.rel_l2(got, ref) < 1e-4 * 0.05, "complex STFT far tighter than the 1e-4 budget expected"
    # leading batch dims are packed like torchaudio does
    got2 = T0(x.to(dev).reshape(3, 1, L)).cpu()
    assert got2.shape == (3, 1) + ref.shape[1:]
    assert torch.equal(got2.reshape(ref.shape), got)


def test_stft_general_hop(dev):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    x, _ = synth.make_batch(2, 3000, 16000, start=20)
    for n_fft, hop in [(512, 128), (400, 100), (1024, 300)]:
        ref = dsp.stft(x, n_fft, hop)
        got = adb.Spectrogram(power=None, n_fft=n_fft, hop_length=hop).to(dev)(x.to(dev)).cpu()
        assert got.shape == ref.shape
        assert metrics.rel_l2(got, ref) < 5e-6


# ---------------------------------------------------------------------------------------------- K1+K2
@pytest.mark.parametrize("n_fft,hop,sr", [(1024, 512, 16000), (640, 320, 16000), (1536, 768, 48000), (1024, 512, 48000)])
def test_logmel_matches_oracle(dev, n_fft, hop, sr):
    import audio_denoising_b200 as adb
    from audio_denoising_b200 import _cabi, _runtime

    dsp, metrics, *_ , synth = _oracle()
    x, _ = synth.make_batch(4, 6000, sr, start=30)
    fb = dsp.mel_fbanks(n_fft // 2 + 1, 64, sr)
    ref = dsp.log_mel(x, n_fft, hop, fb)  # [B, M, T]
    plan = _runtime.get_plan(n_fft, hop, 64, sr, dev)
    assert torch.equal(plan.fb, fb), "host filterbank must be bit-identical to torchaudio's"
    xd = x.to(dev)
    T = plan.num_frames(x.shape[1])
    bt = torch.empty(4, T, 64, device=dev)
    bm = torch.empty(4, 64, T, device=dev)
    _cabi.check(_cabi.lib().b2d_stft_mel_log1p(plan.handle, xd.data_ptr(), None, 4, x.shape[1], bt.data_ptr(), bm.data_ptr(), None,
                                                torch.cuda.current_stream().cuda_stream))
    assert metrics.rel_l2(bm.cpu(), ref) < 1e-4
    assert metrics.rel_l2(bm.cpu(), ref) < 5e-6
    assert torch.equal(bt.transpose(1, 2).contiguous(), bm)
    # the un-fused drop-ins compose to the same thing
    mag = adb.Spectrogram(power=None, n_fft=n_fft, hop_length=hop).to(dev)(xd).abs()
    mel = adb.MelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr).to(dev)(mag)
    assert metrics.rel_l2(mel.log1p().cpu(), ref) < 5e-6


@pytest.mark.parametrize("n_fft,sr", [(1024, 16000), (640, 16000), (1536, 48000)])
@pytest.mark.parametrize("L", [20000, 20003])
def test_logmel_register_kernels_match_oracle_and_generic(dev, n_fft, sr, L):
    """The fused STFT + Mel + log1p register kernels (gl_fast.cu / gl_reg.cu: batch calls asking for the [B, T, n_mels] layout
    only) against the oracle and against the generic shared-memory kernel; L = 20003 takes the unaligned (non-TMA) staging."""
    from audio_denoising_b200 import _cabi, _runtime

    dsp, metrics, *_ , synth = _oracle()
    hop = n_fft // 2
    B = 5
    x, _ = synth.make_batch(B, L, sr, start=33)
    ref = dsp.log_mel(x, n_fft, hop, dsp.mel_fbanks(n_fft // 2 + 1, 64, sr)).transpose(1, 2)
    xd = x.to(dev)
    peak = (torch.rand(B, device=dev) + 0.5)
    outs = []
    for flags in (0, _runtime.PLAN_GENERIC_KERNELS):
        plan = _runtime.get_plan(n_fft, hop, 64, sr, dev, flags=flags)
        T = plan.num_frames(L)
        bt = torch.empty(B, T, 64, device=dev)
        _cabi.check(_cabi.lib().b2d_stft_mel_log1p(plan.handle, xd.data_ptr(), None, B, L, bt.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream))
        sc = torch.empty_like(bt)
        _cabi.check(_cabi.lib().b2d_stft_mel_log1p(plan.handle, xd.data_ptr(), peak.data_ptr(), B, L, sc.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream))
        outs.append((bt.cpu(), sc.cpu()))
    assert metrics.rel_l2(outs[0][0], ref) < 5e-6
    assert metrics.rel_l2(outs[0][0], outs[1][0]) < 2e-6
    ref_sc = dsp.log_mel(x / peak.cpu()[:, None], n_fft, hop, dsp.mel_fbanks(n_fft // 2 + 1, 64, sr)).transpose(1, 2)
    assert metrics.rel_l2(outs[0][1], ref_sc) < 5e-6 and metrics.rel_l2(outs[1][1], ref_sc) < 5e-6


# ---------------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("n_fft,sr,T", [(1024, 16000, 37), (1536, 48000, 3), (640, 16000, 130)])
def test_inverse_mel_matches_lstsq(dev, n_fft, sr, T):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ = _oracle()
    g = torch.Generator().manual_seed(5)
    mel = torch.rand(3, 64, T, generator=g) * 4
    fb = dsp.mel_fbanks(n_fft // 2 + 1, 64, sr)
    ref = dsp.inverse_mel(mel, fb)
    got = adb.InverseMelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr).to(dev)(mel.to(dev)).cpu()
    assert got.shape == ref.shape
    assert metrics.rel_l2(got, ref) < 2e-5
    with pytest.raises(ValueError):
        adb.InverseMelScale(n_mels=64, n_stft=n_fft // 2 + 1, sample_rate=sr).to(dev)(mel[:, :32].to(dev))


def test_inverse_mel_rank_deficient_raises(dev):
    import audio_denoising_b200 as adb

    with pytest.raises(ValueError):
        adb.InverseMelScale(n_mels=64, n_stft=257, sample_rate=48000).to(dev)(torch.zeros(1, 64, 4, device=dev))


# ---------------------------------------------------------------------------------------------- K7
@pytest.mark.parametrize("n_fft,hop", GEOMS)
def test_istft_matches_oracle_and_round_trips(dev, n_fft, hop):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    L = hop * 13
    x, _ = synth.make_batch(2, L, 16000, start=40)
    spec = dsp.stft(x, n_fft, hop)
    ref = dsp.istft(spec, n_fft, hop)
    I0 = adb.InverseSpectrogram(n_fft=n_fft, win_length=n_fft, hop_length=hop).to(dev)
    got = I0(spec.to(dev)).cpu()
    assert got.shape == ref.shape == (2, L)
    assert metrics.rel_l2(got, ref) < 5e-6
    assert metrics.rel_l2(got, x) < 5e-6  # perfect reconstruction (COLA)
    # polar(mag, angle(spec)) variant used by the server path
    mag = spec.abs() * 0.5 + 0.1
    ref2 = dsp.istft(torch.polar(mag, spec.angle()), n_fft, hop)
    got2 = I0(spec.to(dev), magnitude=mag.to(dev)).cpu()
    assert metrics.rel_l2(got2, ref2) < 2e-5


# ---------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("name", CHECKPOINTS)
def test_model_matches_reference_golden(dev, golden, name):
    _, metrics, model, *_ = _oracle()
    m, sd, cfg = _our_model(name, dev)
    z = golden("model_io.npz")
    x = torch.from_numpy(z["x"]).to(dev)
    h0 = torch.from_numpy(z["h0"]).to(dev)
    y, h = m(x)
    assert metrics.rel_l2(y.cpu(), torch.from_numpy(z[f"{name}_y"])) < 1e-5
    assert metrics.rel_l2(h.cpu(), torch.from_numpy(z[f"{name}_h"])) < 1e-5
    y2, h2 = m(x, h0)
    assert metrics.rel_l2(y2.cpu(), torch.from_numpy(z[f"{name}_y_h0"])) < 1e-5
    assert metrics.rel_l2(h2.cpu(), torch.from_numpy(z[f"{name}_h_h0"])) < 1e-5
    assert torch.equal(h0.cpu(), torch.from_numpy(z["h0"])), "caller's hx must not be mutated"
    # chunked carry == full sequence
    ya, ha = m(x[:, :3])
    yb, hb = m(x[:, 3:], ha)
    assert torch.equal(torch.cat([ya, yb], 1), y) and torch.equal(hb, h)
    # 2-D input path (gruunet2.py:291-293)
    y2d, h2d = m(x[0])
    assert y2d.shape == (x.shape[1], 64) and h2d.shape == (1, 17, 4)
    assert metrics.rel_l2(y2d.cpu(), torch.from_numpy(z[f"{name}_y2d"])) < 1e-5


def test_model_random_weights_long_sequence(dev):
    import audio_denoising_b200 as adb

    _, metrics, model, *_ = _oracle()
    sd = model.random_state_dict(seed=3)
    m = adb.GRUUNet2(**model.default_config())
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    orc = model.GRUUNet2Oracle(sd)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(5, 40, 64, generator=g) * 3
    ry, rh = orc(x)
    y, h = m(x.to(dev))
    assert metrics.rel_l2(y.cpu(), ry) < 1e-5
    assert metrics.rel_l2(h.cpu(), rh) < 1e-5


def test_model_rejects_cpu_and_bad_shapes(dev):
    m, *_ = _our_model("good", dev)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 32, device=dev))


# ---------------------------------------------------------------------------------------------- K6
@pytest.mark.parametrize("n_fft,hop,L,B", [(1024, 512, 16000, 3), (512, 256, 4096, 2), (640, 320, 6400, 2), (1536, 768, 1536, 4), (2048, 1024, 20000, 1), (1024, 512, 64000, 2),
                                            (640, 320, 32000, 2), (1536, 768, 48000, 2), (1536, 768, 10000, 3)])
def test_griffinlim_matches_oracle(dev, n_fft, hop, L, B):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    x, _ = synth.make_batch(B, L, 16000, start=50)
    mag = dsp.stft(x, n_fft, hop).abs()
    init = synth.gl_init_angles(mag.shape, seed=7)
    ref = dsp.griffinlim(mag, n_fft, hop, 32, 0.99, init)
    gl = adb.GriffinLim(n_fft=n_fft, win_length=n_fft, hop_length=hop, window_fn=torch.hann_window, power=1.0).to(dev)
    got = gl(mag.to(dev), init_angles=init.to(dev)).cpu()
    assert got.shape == ref.shape
    sdr = metrics.si_sdr(got, ref)
    # Two iterations prove the arithmetic (before Griffin-Lim's own sensitivity enters): > 80 dB on every clip
    # (measured 85-125 dB; an indexing or window bug gives < 20 dB).
    gl2 = adb.GriffinLim(n_fft=n_fft, win_length=n_fft, hop_length=hop, window_fn=torch.hann_window, power=1.0, n_iter=2).to(dev)
    sdr2 = metrics.si_sdr(gl2(mag.to(dev), init_angles=init.to(dev)).cpu(), dsp.griffinlim(mag, n_fft, hop, 2, 0.99, init))
    assert sdr2.min() >= 80.0, f"2 iterations: {sdr2.tolist()} dB"
    # 32 iterations: >= 60 dB for a well-conditioned clip.  Some clips are not: a bin whose rebuilt value nearly
    # cancels flips its direction under the unit-modulus projection and the oracle then disagrees with ITSELF at 45-60 dB when
    # its input moves by 1e-6 (profiles/r2_gl_conditioning.txt); the generic kernels land within 1 dB of the fast ones on those
    # clips (measured: n_fft 640, L 32000, clip 0: 42.8 / 42.4 dB, > 100 dB after two iterations).  So: the typical clip at the
    # survey's bar, no clip below 40 dB.
    assert sdr.max() >= 60.0 and sdr.min() >= 40.0, f"SI-SDR(ours, oracle) = {sdr.tolist()} dB"
    # spectral convergence within 1% of the oracle's
    sc_ref = metrics.rel_l2(dsp.stft(ref, n_fft, hop).abs(), mag)
    sc_got = metrics.rel_l2(dsp.stft(got, n_fft, hop).abs(), mag)
    assert abs(sc_got - sc_ref) <= 0.01 * max(sc_ref, 1e-6) + 1e-6


def test_griffinlim_few_iterations_and_ones_init(dev):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    n_fft, hop = 1024, 512
    x, _ = synth.make_batch(2, 8192, 16000, start=60)
    mag = dsp.stft(x, n_fft, hop).abs()
    for n_iter, mom in [(0, 0.99), (1, 0.99), (2, 0.99), (5, 0.0)]:
        ref = dsp.griffinlim(mag, n_fft, hop, n_iter, mom, None, rand_init=False)
        gl = adb.GriffinLim(n_fft=n_fft, hop_length=hop, power=1.0, n_iter=n_iter, momentum=mom, rand_init=False).to(dev)
        got = gl(mag.to(dev)).cpu()
        assert metrics.si_sdr(got, ref).min() >= 80.0, (n_iter, mom)
    with pytest.raises(ValueError):
        adb.GriffinLim(n_fft=n_fft, hop_length=hop, momentum=1.0)


def test_griffinlim_rand_init_runs_and_converges(dev):
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    x, _ = synth.make_batch(2, 16000, 16000, start=70)
    mag = dsp.stft(x, 1024, 512).abs()
    got = adb.GriffinLim(n_fft=1024, hop_length=512, power=1.0).to(dev)(mag.to(dev)).cpu()
    assert torch.isfinite(got).all()
    assert metrics.rel_l2(dsp.stft(got, 1024, 512).abs(), mag) < 0.35


# ---------------------------------------------------------------------------------------------- chain
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_pipeline_matches_reference_golden(dev, golden, tag):
    """End to end against outputs of the reference's own torchaudio + gruunet2 path (tests/golden/make_golden.py)."""
    import audio_denoising_b200 as adb

    dsp, metrics, *_ = _oracle()
    c = golden("dsp_chain.npz")
    n_fft, hop, sr, L = [int(v) for v in c[f"{tag}_cfg"]]
    m, *_ = _our_model("good", dev)
    pipe = adb.DenoisePipeline(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=sr, n_iter=32)
    noisy = torch.from_numpy(c[f"{tag}_noisy"])
    r = pipe.denoise(noisy.to(dev), init_angles=torch.from_numpy(c[f"{tag}_init"]).to(dev), normalise=False, return_intermediates=True)
    assert metrics.rel_l2(r["logmel"].cpu(), torch.from_numpy(c[f"{tag}_logmel"])) < 1e-4
    assert metrics.rel_l2(r["pred"].cpu(), torch.from_numpy(c[f"{tag}_pred"])) < 2e-5
    assert metrics.rel_l2(r["lin_mag"].cpu(), torch.from_numpy(c[f"{tag}_lin"])) < 5e-5
    sdr = metrics.si_sdr(r["wave"].cpu(), torch.from_numpy(c[f"{tag}_wave"]))
    assert sdr.min() >= 50.0, f"SI-SDR vs reference waveform {sdr.tolist()}"


def test_pipeline_sisdr_vs_clean_within_budget(dev):
    """north_star: denoised waveform SI-SDR within 0.05 dB of the reference path's (same init)."""
    import audio_denoising_b200 as adb

    dsp, metrics, model, pipeline, synth = _oracle()
    noisy, clean = synth.make_batch(4, 16000, 16000, start=80)
    for name in ["good", "dari_tult2"]:
        m, sd, cfg = _our_model(name, dev)
        orc = model.GRUUNet2Oracle(sd, cfg)
        T = 1 + 16000 // 512
        init = synth.gl_init_angles((4, 513, T), seed=7)
        ref = pipeline.denoise_batch(noisy, orc, 1024, 512, 64, 16000, 32, 0.99, init)
        pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000)
        wave, hx = pipe.denoise(noisy.to(dev), init_angles=init.to(dev))
        Lout = wave.shape[1]
        a = metrics.si_sdr(wave.cpu(), clean[:, :Lout])
        b = metrics.si_sdr(ref["wave"], clean[:, :Lout])
        assert (a - b).abs().max() <= 0.05, f"{name}: ours {a.tolist()} vs oracle {b.tolist()}"
        assert metrics.rel_l2(hx.cpu(), ref["hx"]) < 1e-5


def test_noisy_phase_chain_matches_reference_golden(dev, golden):
    import audio_denoising_b200 as adb

    _, metrics, *_ = _oracle()
    s = golden("server_chain.npz")
    m, *_ = _our_model("good", dev)
    pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=48000)
    x = torch.from_numpy(s["x"]).to(dev)
    hx = None
    for rep in range(2):
        wave, hx = pipe.denoise_noisy_phase(x, hx)
        assert metrics.si_sdr(wave.cpu(), torch.from_numpy(s[f"wave{rep}"])).min() >= 70.0
        assert metrics.rel_l2(hx.cpu(), torch.from_numpy(s[f"hx{rep}"])) < 1e-5


@pytest.mark.parametrize("tag", ["s16k", "s48k"])
def test_streaming_matches_reference_recv_golden(dev, golden, tag):
    """app3.DenoisingAudioProcessor.recv (run under stubs by make_golden.py) hop by hop."""
    import audio_denoising_b200 as adb

    _, metrics, *_ = _oracle()
    s = golden("stream.npz")
    n_fft, hop, sr, ncalls = [int(v) for v in s[f"{tag}_cfg"]]
    m, *_ = _our_model("dari_tult2", dev)

    def angles(i, shape):
        # recv call j produces hop j-1 (first window is complete after 2 hop-sized chunks)
        torch.manual_seed(1000 + i + 1)
        return torch.rand(shape, dtype=torch.complex64)

    sd = adb.StreamingDenoiser(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=sr, sessions=1, angles_fn=angles)
    pcm = s[f"{tag}_pcm"]
    ref = s[f"{tag}_out"].astype(np.float64)
    for i in range(ncalls):
        out = sd.push(pcm[i * hop : (i + 1) * hop])
        if out.shape[1] == 0:
            continue  # reference returns a passthrough frame here (app3.py:228-241)
        got = adb.StreamingDenoiser.to_int16(out[0]).astype(np.float64)
        err = np.abs(got - ref[i]).max()
        assert err <= 3.0, f"hop {i}: max int16 deviation {err}"
    assert metrics.rel_l2(sd.hx.cpu(), torch.from_numpy(s[f"{tag}_hx"])) < 1e-4


# ---------------------------------------------------------------------------------------------- full size
def test_full_size_properties_config2(dev):
    """BASELINE config 2 geometry (256 x 4 s @ 16 kHz): size-independent properties at full size."""
    import audio_denoising_b200 as adb

    dsp, metrics, *_ , synth = _oracle()
    B, L, n_fft, hop = 256, 64000, 1024, 512
    x = synth.make_batch_fast(B, L).to(dev)
    T0 = adb.Spectrogram(power=None, n_fft=n_fft, hop_length=hop).to(dev)
    I0 = adb.InverseSpectrogram(n_fft=n_fft, hop_length=hop).to(dev)
    spec = T0(x)
    assert spec.shape == (B, 513, 126)
    back = I0(spec)
    assert back.shape == (B, 64000)
    assert metrics.rel_l2(back[::37].cpu(), x[::37].cpu()) < 5e-6  # stft -> istft identity
    # linearity of the STFT
    s2 = T0(0.5 * x[:8] + 0.25 * x[8:16])
    assert metrics.rel_l2(s2.cpu(), (0.5 * spec[:8] + 0.25 * spec[8:16]).cpu()) < 1e-5
    # Griffin-Lim on consistent magnitudes: batch entries are independent (same clip twice -> same answer)
    mag = spec.abs()
    init = torch.rand((2, 513, 126), dtype=torch.complex64, device=dev)
    gl = adb.GriffinLim(n_fft=n_fft, hop_length=hop, power=1.0).to(dev)
    big = gl(mag[:64], init_angles=init[:1].expand(64, -1, -1).contiguous())
    small = gl(mag[5:6], init_angles=init[:1])
    assert metrics.si_sdr(big[5:6].cpu(), small.cpu()).min() >= 70.0  # same answer whatever the batch / run partition (up to fp32 rounding amplified by 32 iterations)
    # spot-check a few clips of the full batch against the CPU oracle
    ref = dsp.griffinlim(mag[[0, 63]].cpu(), n_fft, hop, 32, 0.99, init[:1].expand(2, -1, -1).cpu())
    assert metrics.si_sdr(big[[0, 63]].cpu(), ref).min() >= 60.0


# ---------------------------------------------------------------------------------------------- fast vs generic kernels
@pytest.mark.parametrize("n_fft,hop", [(1024, 512), (512, 256), (2048, 1024)])
def test_fast_kernels_match_generic_kernels(dev, n_fft, hop):
    """n_fft 1024 / 512 / 2048 have register/TMA fast kernels (gl_fast*.cu); the generic shared-memory kernels (plan flag
    B2D_PLAN_GENERIC_KERNELS) are the cross-check.  Same C-ABI calls, same seed for the in-kernel rand_init draws."""
    import audio_denoising_b200 as adb
    from audio_denoising_b200 import _cabi, _runtime

    dsp, metrics, *_ , synth = _oracle()
    B, L = 5, 20000
    x, _ = synth.make_batch(B, L, 16000, start=90)
    xd = x.to(dev)
    plan_fast = _runtime.get_plan(n_fft, hop, 64, 16000, dev)
    plan_generic = _runtime.get_plan(n_fft, hop, 64, 16000, dev, flags=_runtime.PLAN_GENERIC_KERNELS)
    T = plan_fast.num_frames(L)
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream

    def logmel(plan):
        out = torch.empty(B, T, 64, device=dev)
        _cabi.check(lib.b2d_stft_mel_log1p(plan.handle, xd.data_ptr(), None, B, L, out.data_ptr(), None, None, st))
        return out.cpu()

    def gl(plan, seed, n_iter):
        mag = dsp.stft(x, n_fft, hop).abs().to(dev)
        ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(plan.handle, B, T), dtype=torch.uint8, device=dev)
        wave = torch.empty(B, plan.out_length(T), device=dev)
        _cabi.check(lib.b2d_griffinlim(plan.handle, mag.data_ptr(), None, seed, B, T, n_iter, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st))
        return wave.cpu()

    fast, slow = (dict(logmel=logmel(pl), gl0=gl(pl, 77, 0), gl4=gl(pl, 77, 4), gl32=gl(pl, 77, 32), ones=gl(pl, 0, 8))
                  for pl in (plan_fast, plan_generic))
    assert metrics.rel_l2(fast["logmel"], slow["logmel"]) < 2e-6
    assert metrics.rel_l2(fast["logmel"], dsp.log_mel(x, n_fft, hop, dsp.mel_fbanks(n_fft // 2 + 1, 64, 16000)).transpose(1, 2)) < 5e-6
    assert metrics.si_sdr(fast["gl0"], slow["gl0"]).min() > 110.0  # same random initial phase from the same seed
    sdr4 = metrics.si_sdr(fast["gl4"], slow["gl4"])
    assert sdr4.median() > 100.0 and sdr4.min() > 80.0, sdr4.tolist()
    # 32 iterations from a random phase amplify last-bit differences where the spectrum is nearly empty (measured: one
    # clip in five drops to 38-48 dB in a few hop-blocks -- which one depends on the run partition, i.e. on the summation
    # order at run boundaries -- while all agree to > 100 dB after 4 iterations): bound the typical clip tightly, the
    # worst clip loosely, and require the same spectral convergence
    sdr32 = metrics.si_sdr(fast["gl32"], slow["gl32"])
    assert sdr32.median() > 60.0 and sdr32.min() > 30.0, sdr32.tolist()
    mag_ref = dsp.stft(x, n_fft, hop).abs()
    for i in range(B):
        sc_f = metrics.rel_l2(dsp.stft(fast["gl32"][i : i + 1], n_fft, hop).abs(), mag_ref[i : i + 1])
        sc_s = metrics.rel_l2(dsp.stft(slow["gl32"][i : i + 1], n_fft, hop).abs(), mag_ref[i : i + 1])
        assert abs(sc_f - sc_s) <= 0.02 * sc_s + 1e-6, (i, sc_f, sc_s)
    assert metrics.si_sdr(fast["ones"], slow["ones"]).min() > 90.0
    assert metrics.si_sdr(fast["ones"], dsp.griffinlim(dsp.stft(x, n_fft, hop).abs(), n_fft, hop, 8, 0.99, None, rand_init=False)).min() > 80.0
    # a different seed gives a different (but equally consistent) reconstruction
    other = gl(plan_fast, 78, 32)
    assert metrics.si_sdr(other, fast["gl32"]).max() < 30.0
    mag = dsp.stft(x, n_fft, hop).abs()
    assert metrics.rel_l2(dsp.stft(other, n_fft, hop).abs(), mag) < 0.35


@pytest.mark.parametrize("n_fft", [512, 640, 1024, 1536])
def test_single_launch_griffinlim_kernels_agree(dev, n_fft):
    """The single-launch Griffin-Lim kernels for small problems (csrc/gl_reg.cu): the cooperative whole-GPU kernel (default) against
    the cluster-per-clip kernel (plan flag CLUSTER_GL) -- same frame arithmetic, same two-slots-per-frame format: BIT-IDENTICAL --
    and against the per-iteration batch kernels where no cluster plan exists (different association of the overlap-add sums:
    > 100 dB after a few iterations), seeded in-kernel random initial phase and injected initial angles."""
    from audio_denoising_b200 import _cabi, _runtime

    _, metrics, *_ = _oracle()
    hop = n_fft // 2
    plan_coop = _runtime.get_plan(n_fft, hop, 0, 0, dev)
    plan_cluster = _runtime.get_plan(n_fft, hop, 0, 0, dev, flags=_runtime.PLAN_CLUSTER_GL)
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(n_fft + 1)

    def run(pl, mag, B, T, n_iter, seed, ang=None):
        ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(pl.handle, B, T), dtype=torch.uint8, device=dev)
        wave = torch.zeros(B, pl.out_length(T), device=dev)
        _cabi.check(lib.b2d_griffinlim_frames(pl.handle, mag.data_ptr(), None if ang is None else ang.data_ptr(), seed, B, T, n_iter, 0.99,
                                              None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st))
        return wave.cpu()

    for B, T, exact in [(1, 126, True), (3, 40, True), (2, 9, True), (5, 5, True), (16, 126, False), (24, 31, False)]:
        mag = (torch.rand(B, T, plan_coop.frame_stride, generator=g) * 2).to(dev)
        for n_iter, seed in [(0, 11), (1, 11), (5, 12), (4, 0)]:
            a, b = run(plan_coop, mag, B, T, n_iter, seed), run(plan_cluster, mag, B, T, n_iter, seed)
            assert torch.isfinite(a).all()
            if exact:
                assert torch.equal(a, b), (n_fft, B, T, n_iter, seed, float(metrics.si_sdr(a, b).min()))
            else:
                sdr = metrics.si_sdr(a, b)
                assert sdr.median() > 100.0 and sdr.min() > 40.0, (n_fft, B, T, n_iter, float(sdr.median()), float(sdr.min()))
        ang = torch.polar(torch.ones(B, n_fft // 2 + 1, T), torch.rand(B, n_fft // 2 + 1, T, generator=g) * 6.28).to(dev).contiguous()
        a, b = run(plan_coop, mag, B, T, 3, 0, ang), run(plan_cluster, mag, B, T, 3, 0, ang)
        if exact:
            assert torch.equal(a, b), (n_fft, B, T, "angles")
        else:
            assert metrics.si_sdr(a, b).median() > 100.0


def test_rand_init_draws_are_uniform(dev):
    """rand_init=True statistics: with mag == 1 and zero iterations the output is istft(angles_0); its STFT on
    interior frames recovers the projection of the draws, whose mean must match U[0,1) (0.5) closely."""
    import audio_denoising_b200 as adb
    from audio_denoising_b200 import _cabi, _runtime

    plan = _runtime.get_plan(1024, 512, 0, 0, dev)
    B, T = 4, 40
    lib = _cabi.lib()
    mag = torch.zeros(B, 513, T, device=dev)
    mag[:, 0, :] = 1.0  # only the DC bin: frame t contributes Re(angle_0[0, t]) / N * window
    ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(plan.handle, B, T), dtype=torch.uint8, device=dev)
    wave = torch.empty(B, plan.out_length(T), device=dev)
    _cabi.check(lib.b2d_griffinlim(plan.handle, mag.data_ptr(), None, 4242, B, T, 0, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(),
                                   torch.cuda.current_stream().cuda_stream))
    # sample at the centre of frame t (window == 1, neighbour windows == 0): x = Re(a_t) / N
    centres = wave[:, 511::512][:, : T - 2] * 1024.0
    assert 0.0 <= float(centres.min()) and float(centres.max()) < 1.0
    assert abs(float(centres.mean()) - 0.5) < 0.08
    assert float(centres.std()) > 0.2


# ---------------------------------------------------------------------------------------------- tcgen05 path
@pytest.mark.parametrize("mode,tol", [("mma", 1e-5), ("utc", 1e-5), ("fp32", 1e-5)])
def test_model_tensor_core_modes(dev, golden, mode, tol):
    """The three convolution engines against the reference goldens: mma = warp-level m16n8k8 TF32 MMAs with the big + small split
    (unet_mma.cu, default), utc = the encoder as one persistent tcgen05 / TMEM kernel (unet_tc.cu), fp32 = CUDA-core FMA (model.cu)."""
    _, metrics, model, *_ = _oracle()
    z = golden("model_io.npz")
    x = torch.from_numpy(z["x"]).to(dev)
    for name in CHECKPOINTS:
        m, sd, cfg = _our_model(name, dev)
        m.conv_mode = mode
        y, h = m(x)
        assert metrics.rel_l2(y.cpu(), torch.from_numpy(z[f"{name}_y"])) < tol, name
        assert metrics.rel_l2(h.cpu(), torch.from_numpy(z[f"{name}_h"])) < tol, name
    # a batch that is not a multiple of the 128-row tiles, long sequence
    sd = model.random_state_dict(seed=5)
    import audio_denoising_b200 as adb

    m = adb.GRUUNet2(**model.default_config())
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    m.conv_mode = mode
    g = torch.Generator().manual_seed(11)
    xx = torch.rand(3, 37, 64, generator=g) * 3
    ry, rh = model.GRUUNet2Oracle(sd)(xx)
    y, h = m(xx.to(dev))
    assert metrics.rel_l2(y.cpu(), ry) < tol and metrics.rel_l2(h.cpu(), rh) < tol


@pytest.mark.parametrize("mode", ["utc", "mma"])
def test_pipeline_tensor_core_mode_within_budget(dev, mode):
    """Whole chain with conv_mode utc / mma (incl. the tcgen05 inverse-mel GEMM): same 0.05 dB SI-SDR budget."""
    import audio_denoising_b200 as adb

    dsp, metrics, model, pipeline, synth = _oracle()
    noisy, clean = synth.make_batch(4, 16000, 16000, start=80)
    m, sd, cfg = _our_model("good", dev)
    m.conv_mode = mode
    T = 1 + 16000 // 512
    init = synth.gl_init_angles((4, 513, T), seed=7)
    ref = pipeline.denoise_batch(noisy, model.GRUUNet2Oracle(sd, cfg), 1024, 512, 64, 16000, 32, 0.99, init)
    pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000)
    r = pipe.denoise(noisy.to(dev), init_angles=init.to(dev), return_intermediates=True)
    assert metrics.rel_l2(r["pred"].cpu(), ref["pred"]) < 1e-5
    assert metrics.rel_l2(r["lin_mag"].cpu(), ref["lin_mag"]) < 2e-5
    Lout = r["wave"].shape[1]
    a = metrics.si_sdr(r["wave"].cpu(), clean[:, :Lout])
    b = metrics.si_sdr(ref["wave"], clean[:, :Lout])
    assert (a - b).abs().max() <= 0.05


# ---------------------------------------------------------------------------------------------- server loop (section 8f rank 1)
def test_server_request_handler_and_wire_protocol(dev, golden):
    """DenoiseServer.handle == server.py:199-220 on [n, channels] arrays; then the same through a real Listener/Client pair."""
    import threading
    from multiprocessing.connection import Client

    import audio_denoising_b200 as adb

    _, metrics, *_ = _oracle()
    s = golden("server_chain.npz")
    m, *_ = _our_model("good", dev)
    srv = adb.DenoiseServer(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=48000)
    X = np.stack([s["x"][0], -s["x"][0]], axis=1)  # [n, 2]: only channel 0 is used, answer is repeated over channels
    for rep in range(2):
        O = srv.handle(X)
        assert O.shape == (s[f"wave{rep}"].shape[1], 2) and np.array_equal(O[:, 0], O[:, 1])
        assert metrics.si_sdr(torch.from_numpy(O[:, 0]), torch.from_numpy(s[f"wave{rep}"][0])) >= 70.0
    assert metrics.rel_l2(srv.hx.cpu(), torch.from_numpy(s["hx1"])) < 1e-5
    # wire protocol: pickled ndarray in, ndarray out, hx carried across requests, 'close' ends the connection
    srv2 = adb.DenoiseServer(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=48000)
    address = ("localhost", 6117)
    t = threading.Thread(target=srv2.serve_forever, kwargs=dict(address=address, max_requests=2), daemon=True)
    t.start()
    conn = None
    for _ in range(100):
        try:
            conn = Client(address)
            break
        except ConnectionRefusedError:
            import time

            time.sleep(0.05)
    assert conn is not None
    outs = []
    for rep in range(2):
        conn.send(s["x"][0].reshape(-1, 1))
        outs.append(conn.recv())
    conn.close()
    t.join(timeout=20)
    for rep in range(2):
        assert metrics.si_sdr(torch.from_numpy(outs[rep][:, 0]), torch.from_numpy(s[f"wave{rep}"][0])) >= 70.0


def test_streaming_cuda_graph_matches_eager_launches(dev):
    """The captured per-hop CUDA graph (seed read from device memory) replays exactly what the eager launches compute."""
    import audio_denoising_b200 as adb

    m, *_ = _our_model("good", dev)
    rng = np.random.default_rng(3)
    sig = (rng.standard_normal((2, 640 + 320 * 6)) * 0.2).astype(np.float32)
    outs = {}
    for use_graph in (True, False):
        torch.manual_seed(5)
        sd = adb.StreamingDenoiser(m, n_fft=640, hop_length=320, n_mels=64, sample_rate=16000, sessions=2, use_graph=use_graph)
        outs[use_graph] = (sd.push(sig), sd.hx.clone(), sd.ola.clone())
    assert outs[True][0].shape == (2, 320 * 7)  # 7 full windows in 640 + 6 * 320 samples
    assert np.array_equal(outs[True][0], outs[False][0])
    assert torch.equal(outs[True][1], outs[False][1]) and torch.equal(outs[True][2], outs[False][2])
    assert np.abs(outs[True][0][:, 320:]).max() > 0  # something was emitted after the one-hop delay


# ---------------------------------------------------------------------------------------------- ingest (section 8f rank 2)
@pytest.mark.parametrize("tag", ["s16k", "s48k"])
def test_recv_pcm_matches_reference_recv_including_passthrough(dev, golden, tag):
    """Every recv call of app3.py:167-250, including the pass-through frames of app3.py:228-241 (bit-exact)."""
    import audio_denoising_b200 as adb

    s = golden("stream.npz")
    n_fft, hop, sr, ncalls = [int(v) for v in s[f"{tag}_cfg"]]
    m, *_ = _our_model("dari_tult2", dev)

    def angles(i, shape):
        torch.manual_seed(1000 + i + 1)
        return torch.rand(shape, dtype=torch.complex64)

    sd = adb.StreamingDenoiser(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=sr, sessions=1, angles_fn=angles)
    pcm, ref = s[f"{tag}_pcm"], s[f"{tag}_out"]
    passthrough = 0
    for i in range(ncalls):
        chunk = pcm[i * hop : (i + 1) * hop]
        got = sd.recv_pcm(chunk.reshape(-1, 1))
        assert got.dtype == np.int16 and got.shape == (hop,)
        if np.array_equal(ref[i], chunk):
            passthrough += 1
            assert np.array_equal(got, ref[i]), f"call {i}: pass-through frame differs"
        else:
            assert np.abs(got.astype(np.int64) - ref[i].astype(np.int64)).max() <= 3
    assert passthrough >= 1


@pytest.mark.parametrize("shape,channel", [((4801,), 0), ((4801, 2), 0), ((4801, 2), 1), ((4801, 2), -1), ((0, 1), 0)])
def test_pcm16_to_float_bit_exact(dev, shape, channel):
    import audio_denoising_b200 as adb

    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32768, size=shape, dtype=np.int16)
    got = adb.pcm16_to_float(torch.from_numpy(pcm).to(dev), channel).cpu().numpy()
    a = pcm.reshape(shape[0], shape[1] if len(shape) > 1 else 1).astype(np.float32) / np.float32(32767)
    want = a.mean(axis=1, dtype=np.float32) if (channel < 0 and a.shape[1] > 1) else a[:, max(channel, 0) if a.shape[1] > 1 else 0]
    if channel < 0 and a.shape[1] > 1:
        assert np.abs(got - want).max() <= 1e-7
    else:
        assert np.array_equal(got, want)


def test_float_to_pcm16_bit_exact(dev):
    import audio_denoising_b200 as adb

    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-1.5, 1.5, 100000).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, -1.0, 0.99999, -0.99999, 1e-9, 3.0, -3.0, 0.5 / 32767, -0.5 / 32767], np.float32)])
    want = (np.clip(x, -1.0, 1.0) * 32767).astype(np.int16)
    got = adb.float_to_pcm16(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert np.array_equal(got, want)
    # and the round trip int16 -> float -> int16 is the identity on [-32767, 32767]
    p = torch.arange(-32767, 32768, dtype=torch.int16, device=dev)
    assert torch.equal(adb.float_to_pcm16(adb.pcm16_to_float(p)), p)


@pytest.mark.parametrize("orig,new,L", [(44100, 48000, 44100), (48000, 44100, 48000), (16000, 48000, 5000), (48000, 16000, 12001),
                                        (44100, 48000, 1), (44100, 48000, 146), (8000, 8000, 100)])
def test_resample_matches_torchaudio(dev, orig, new, L):
    """utils.R1 / R2 (utils.py:48-49): same polyphase table (bit-identical, checked on CPU) and the same FIR sums."""
    import torchaudio
    import audio_denoising_b200 as adb

    g = torch.Generator().manual_seed(orig + new + L)
    x = torch.randn(3, L, generator=g) * 0.3
    want = torchaudio.transforms.Resample(orig, new)(x)
    got = adb.Resample(orig, new)(x.to(dev)).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))
    got1 = adb.Resample(orig, new)(x[0].to(dev)).cpu()  # 1-D input keeps its rank
    assert got1.shape == want[0].shape and torch.equal(got1, got[0])


def test_resample_round_trip_44k1_48k(dev):
    """R2(R1(x)) ~ x for band-limited audio (the way dataset clips travel through utils.R1/R2)."""
    from audio_denoising_b200 import utils

    t = torch.arange(44100, dtype=torch.float64) / 44100
    x = (0.4 * torch.sin(2 * np.pi * 440 * t) + 0.2 * torch.sin(2 * np.pi * 3000 * t + 1.0)).float().to(dev)
    y = utils.R2(utils.R1(x))
    assert y.shape[-1] == 44100
    mid = slice(400, 43700)  # away from the zero-padded edges
    assert (y[mid] - x[mid]).abs().max() < 2e-3


@pytest.mark.parametrize("B,L", [(3, 64000), (2, 64001), (5, 7), (1, 16384), (4, 16390), (2, 1)])
def test_peak_is_exact(dev, B, L):
    """K0 (app3.py:181-186): max |x| per clip, exact for every alignment / tail length; silence maps to 1."""
    from audio_denoising_b200 import _cabi

    g = torch.Generator().manual_seed(B * 1000 + L)
    x = torch.randn(B, L, generator=g)
    x[0] *= 0.0  # digital silence -> peak 1 (no division by zero downstream)
    if B > 1:
        x[1, L - 1] = -7.5  # the maximum sits in the last (possibly unaligned) sample
    xd = x.to(dev)
    peak = torch.empty(B, device=dev)
    _cabi.check(_cabi.lib().b2d_peak(xd.data_ptr(), B, L, peak.data_ptr(), torch.cuda.current_stream().cuda_stream))
    want = x.abs().amax(dim=1)
    want = torch.where(want > 1e-6, want, torch.ones_like(want))
    assert torch.equal(peak.cpu(), want)


def test_denoise_host_pcm16_matches_float_path(dev):
    """Host-to-host with int16 PCM on the link == int16 conversion around the float32 path (same injected initial phase)."""
    import audio_denoising_b200 as adb

    *_, synth = _oracle()
    noisy, _ = synth.make_batch(3, 16000, 16000, start=120)
    noisy = noisy / noisy.abs().amax(dim=1, keepdim=True) * 0.9
    m, *_ = _our_model("good", dev)
    pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000)
    T = 1 + 16000 // 512
    init = synth.gl_init_angles((3, 513, T), seed=3).to(dev)
    pcm = (noisy * 32767).to(torch.int16)
    got = pipe.denoise_host(pcm.pin_memory(), init_angles=init)
    assert got.dtype == torch.int16 and got.shape == (3, 512 * (T - 1))
    ref = pipe.denoise_host((pcm.float() / 32767).pin_memory(), init_angles=init)
    want = (ref.clamp(-1, 1) * 32767).to(torch.int16)
    assert (got.int() - want.int()).abs().max() <= 1
    with pytest.raises(TypeError):
        pipe.denoise_host(torch.zeros(2, 16000, dtype=torch.float64).pin_memory())


@pytest.mark.parametrize("depth", [2, 3, 4])
def test_denoise_host_pipelined_batches_keep_their_own_results(dev, depth):
    """Seven batches enqueued back to back with wait=False through a staging ring of 2 / 3 / 4 slots (3 is the default): every
    batch's result equals what the same batch gives alone (same injected initial phase) -- no slot is reused before its kernels
    and its download are done, in either link format."""
    import audio_denoising_b200 as adb

    *_, synth = _oracle()
    m, *_ = _our_model("good", dev)
    B, L = 4, 8192
    T = 1 + L // 512
    init = synth.gl_init_angles((B, 513, T), seed=11).to(dev)
    g = torch.Generator().manual_seed(100 + depth)
    for dtype in (torch.int16, torch.float32):
        batches = []
        for k in range(7):
            x = torch.randn(B, L, generator=g) * 0.2
            batches.append(((x.clamp(-1, 1) * 32767).to(torch.int16) if dtype == torch.int16 else x).pin_memory())
        solo = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=4)
        want = [solo.denoise_host(b, init_angles=init).clone() for b in batches]
        pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=4)
        pipe.host_ring_depth = depth
        outs = [torch.empty_like(want[0]).pin_memory() for _ in batches]
        for b, o in zip(batches, outs):
            pipe.denoise_host(b, o, init_angles=init, wait=False)
        pipe.host_synchronize()
        for k, (o, w) in enumerate(zip(outs, want)):
            assert torch.equal(o, w), (depth, str(dtype), k)


@pytest.mark.parametrize("n_fft", [512, 1024, 2048, 640, 1536])
def test_fast_paths_match_generic_on_ragged_shapes(dev, n_fft):
    """Run partitions of every flavour (single short run, odd run lengths, a last run of one frame, more runs than warp
    slots) for the register fast paths: 3 iterations from all-ones angles against the generic kernel."""
    from audio_denoising_b200 import _cabi, _runtime

    _, metrics, *_ = _oracle()
    hop = n_fft // 2
    plans = [_runtime.get_plan(n_fft, hop, 0, 0, dev), _runtime.get_plan(n_fft, hop, 0, 0, dev, flags=_runtime.PLAN_GENERIC_KERNELS)]
    plan = plans[0]
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(n_fft)
    pooled = []
    for B, T in [(1, 3), (1, 4), (2, 5), (3, 7), (1, 33), (5, 18), (2, 126), (700, 9), (37, 23)]:
        mag = (torch.rand(B, T, plan.frame_stride, generator=g) * 2).to(dev)
        outs = []
        for pl in plans:
            ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(pl.handle, B, T), dtype=torch.uint8, device=dev)
            wave = torch.zeros(B, pl.out_length(T), device=dev)
            _cabi.check(lib.b2d_griffinlim_frames(pl.handle, mag.data_ptr(), None, 0, B, T, 3, 0.99, None, wave.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), st))
            outs.append(wave.cpu())
        assert torch.isfinite(outs[0]).all()
        sdr = metrics.si_sdr(outs[0], outs[1])
        pooled.append(sdr)
        # random magnitudes are not a consistent spectrogram: where a rebuilt value nearly cancels, the unit-modulus projection
        # is discontinuous and a last-bit difference flips one bin's direction (one flipped bin in a clip = ~40 dB; measured
        # over seeds: 0-4 % of the clips, any partition, any kernel pair): allow that for a few clips per case, require
        # > 100 dB for the typical clip over all cases
        assert sdr.min() > 30.0 and int((sdr <= 90.0).sum()) <= max(1, B // 10), (n_fft, B, T, float(sdr.median()), float(sdr.min()))
    pooled = torch.cat(pooled)
    assert pooled.median() > 100.0 and float((pooled <= 90.0).double().mean()) < 0.06, (n_fft, float(pooled.median()))


@pytest.mark.parametrize("n_fft", [512, 1024, 2048, 640, 1536])
def test_fast_paths_match_generic_on_consistent_spectrograms(dev, n_fft):
    """The same ragged run partitions on CONSISTENT spectrograms -- |STFT| of band-limited signals, where no rebuilt bin nearly
    cancels and the unit-modulus projection is continuous -- with the tight bar on EVERY clip: 3 iterations from the true
    phase, register fast path against the generic kernel."""
    import audio_denoising_b200 as adb
    from audio_denoising_b200 import _cabi, _runtime

    _, metrics, *_ = _oracle()
    hop = n_fft // 2
    plans = [_runtime.get_plan(n_fft, hop, 0, 0, dev), _runtime.get_plan(n_fft, hop, 0, 0, dev, flags=_runtime.PLAN_GENERIC_KERNELS)]
    spec = adb.Spectrogram(n_fft=n_fft, hop_length=hop, power=None).to(dev)
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(7 * n_fft)
    worst = 1e9
    for B, T in [(1, 3), (1, 4), (2, 5), (3, 7), (1, 33), (5, 18), (2, 126), (300, 9), (37, 23)]:
        L = hop * (T - 1)
        t = torch.arange(L, dtype=torch.float64) / 16000.0
        f0 = 80.0 + 300.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)
        x = sum((0.6 / k) * torch.sin(2 * torch.pi * k * f0 * t * (1.0 + 0.05 * torch.sin(2 * torch.pi * 3.0 * t)) + k) for k in range(1, 9))
        x = (x + 0.05 * torch.randn(B, L, generator=g, dtype=torch.float64)).float().to(dev)
        X = spec(x)  # [B, F, T] torch layout
        mag = X.abs().contiguous()
        assert mag.shape == (B, n_fft // 2 + 1, T)
        # start from the true phase: the iterates stay next to a consistent spectrogram, no rebuilt bin comes near zero
        ang = torch.where(mag > 0, X / mag.clamp_min(1e-30), torch.ones_like(X)).contiguous()
        outs = []
        for pl in plans:
            ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(pl.handle, B, T), dtype=torch.uint8, device=dev)
            wave = torch.zeros(B, pl.out_length(T), device=dev)
            _cabi.check(lib.b2d_griffinlim(pl.handle, mag.data_ptr(), ang.data_ptr(), 0, B, T, 3, 0.99, None, wave.data_ptr(), ws.data_ptr(),
                                           ws.numel(), st))
            outs.append(wave.cpu())
        sdr = metrics.si_sdr(outs[0], outs[1])
        assert float(metrics.si_sdr(outs[0], x.cpu()[:, :outs[0].shape[1]]).min()) > 60.0  # and the signal itself comes back
        worst = min(worst, float(sdr.min()))
        assert sdr.min() > 90.0, (n_fft, B, T, float(sdr.median()), float(sdr.min()))
    assert worst > 90.0


@pytest.mark.parametrize("n_fft,L,B", [(512, 16000, 2), (2048, 32000, 2), (1024, 160000, 2), (640, 32000, 2), (1536, 40000, 2)])
def test_pipeline_other_geometries_match_oracle(dev, n_fft, L, B):
    """Whole chain vs the oracle for the other Griffin-Lim fast paths (n_fft 512 / 2048) and for the corpus geometry of
    BASELINE config 4 (10 s clips, T = 313), same injected initial phase."""
    import audio_denoising_b200 as adb

    dsp, metrics, model, pipeline, synth = _oracle()
    hop = n_fft // 2
    noisy, clean = synth.make_batch(B, L, 16000, start=140)
    m, sd, cfg = _our_model("good", dev)
    T = 1 + L // hop
    init = synth.gl_init_angles((B, n_fft // 2 + 1, T), seed=11)
    ref = pipeline.denoise_batch(noisy, model.GRUUNet2Oracle(sd, cfg), n_fft, hop, 64, 16000, 32, 0.99, init)
    pipe = adb.DenoisePipeline(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=16000)
    r = pipe.denoise(noisy.to(dev), init_angles=init.to(dev), return_intermediates=True)
    assert metrics.rel_l2(r["logmel"].cpu(), ref["logmel"]) < 1e-4
    assert metrics.rel_l2(r["pred"].cpu(), ref["pred"]) < 2e-5
    assert metrics.rel_l2(r["lin_mag"].cpu(), ref["lin_mag"]) < 5e-5
    assert r["wave"].shape == ref["wave"].shape
    Lout = r["wave"].shape[1]
    a = metrics.si_sdr(r["wave"].cpu(), clean[:, :Lout])
    b = metrics.si_sdr(ref["wave"], clean[:, :Lout])
    assert (a - b).abs().max() <= 0.05, (a.tolist(), b.tolist())
    assert metrics.si_sdr(r["wave"].cpu(), ref["wave"]).min() >= 40.0


# ---------------------------------------------------------------------------------------------- round 2: host-layer contracts
def test_denoise_pcm16_native_call_equals_float_chain(dev):
    """b2d_denoise_batch_pcm16 (ingest fused with the peak pass, float staging in the workspace) is bit-identical to
    pcm/32767 -> b2d_denoise_batch -> clip*32767: same kernels, same injected initial phase."""
    import audio_denoising_b200 as adb

    *_, synth = _oracle()
    for B, L in [(3, 16000), (2, 16391)]:  # 16391: odd length -> unaligned rows, scalar ingest path
        noisy, _ = synth.make_batch(B, L, 16000, start=150)
        m, *_ = _our_model("good", dev)
        pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000)
        T = 1 + L // 512
        init = synth.gl_init_angles((B, 513, T), seed=5).to(dev)
        pcm = (noisy * 0.8 * 32767).to(torch.int16).to(dev)
        got, hx1 = pipe.denoise_pcm16(pcm, init_angles=init)
        xf = torch.from_numpy(pcm.cpu().numpy().astype(np.float32) / np.float32(32767)).to(dev)  # true division (torch CUDA multiplies by 1/s)
        wave, hx2 = pipe.denoise(xf, init_angles=init)
        want = (wave.clamp(-1, 1) * 32767).to(torch.int16)
        assert got.dtype == torch.int16 and got.shape == want.shape
        assert torch.equal(got, want)
        assert torch.equal(hx1, hx2)
        out = torch.empty_like(got)
        r, _ = pipe.denoise_pcm16(pcm, init_angles=init, out=out)
        assert r.data_ptr() == out.data_ptr() and torch.equal(out, want)


def test_pipeline_validates_hx_out_and_device(dev):
    """a stale hx from another batch size or a too-small / strided `out` must raise, not index out of bounds."""
    import audio_denoising_b200 as adb

    m, *_ = _our_model("good", dev)
    pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=2)
    x = torch.randn(3, 8000, device=dev) * 0.1
    Lout = pipe.out_length(8000)
    with pytest.raises(ValueError):
        pipe.denoise(x, hx=torch.zeros(2, 17, 4, device=dev))
    with pytest.raises(ValueError):
        pipe.denoise_noisy_phase(x, hx=torch.zeros(3, 17, 5, device=dev))
    with pytest.raises(ValueError):
        pipe.denoise(x, out=torch.empty(3, Lout - 1, device=dev))
    with pytest.raises(ValueError):
        pipe.denoise(x, out=torch.empty(3, 2 * Lout, device=dev)[:, ::2])
    with pytest.raises(ValueError):
        pipe.denoise(x, out=torch.empty(3, Lout, device=dev, dtype=torch.float16))
    with pytest.raises(RuntimeError):
        pipe.denoise(x.cpu())
    w, h = pipe.denoise(x, hx=torch.zeros(3, 17, 4, device=dev), out=torch.empty(3, Lout, device=dev))
    assert torch.isfinite(w).all() and h.shape == (3, 17, 4)


def test_forward_is_reentrant_across_threads_and_streams(dev):
    """One GRUUNet2 object shared by several sessions (app3.py:46,449): concurrent forwards on different CUDA streams from
    different threads must not share scratch memory -- every result equals the single-threaded one bit for bit."""
    import threading

    m, *_ = _our_model("good", dev)
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand(4 + i, 40 + 7 * i, 64, generator=g).to(dev) for i in range(4)]
    want = [tuple(t.clone() for t in m(x)) for x in xs]
    torch.cuda.synchronize()
    got = [None] * len(xs)
    errors = []

    def worker(i):
        try:
            s = torch.cuda.Stream(dev)
            with torch.cuda.stream(s):
                for _ in range(20):
                    out, hx = m(xs[i])
                s.synchronize()
            got[i] = (out, hx)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(xs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for (o, h), (wo, wh) in zip(got, want):
        assert torch.equal(o, wo) and torch.equal(h, wh)


def test_streaming_graph_follows_weight_changes(dev):
    """a captured hop must not keep replaying a freed / stale weight pack after load_state_dict on a live model."""
    import audio_denoising_b200 as adb

    sd_a, cfg = load_weights("good")
    sd_b, _ = load_weights("dari_tult2")
    rng = np.random.default_rng(5)
    sig = (rng.standard_normal((1, 640 + 320 * 6)) * 0.1).astype(np.float32)

    def run(states):
        m = adb.GRUUNet2(**cfg)
        m.load_state_dict(states[0])
        m = m.to(dev).eval()
        torch.manual_seed(0)  # the denoiser draws the base of its per-hop seed stream from torch's generator when it is built
        s = adb.StreamingDenoiser(m, n_fft=640, hop_length=320, n_mels=64, sample_rate=16000, n_iter=4, sessions=1,
                                  angles_fn=None, use_graph=True)
        outs = []
        for i in range(6):
            if i == 3 and len(states) > 1:
                m.load_state_dict(states[1])
            outs.append(s.step(sig[:, i * 320: i * 320 + 640]))
        return np.concatenate(outs, axis=1), m

    switched, m = run([sd_a, sd_b])
    only_a, _ = run([sd_a])
    assert np.array_equal(switched[:, : 3 * 320], only_a[:, : 3 * 320])  # same seeds, same weights for the first hops
    assert not np.array_equal(switched[:, 3 * 320:], only_a[:, 3 * 320:])  # the new weights took effect in the graph
    # explicit repack() after an edit autograd cannot see
    h0 = m.native_model(dev)
    assert m.native_model(dev) is h0
    with torch.no_grad():
        m.cell.input_gate.downs[0].conv.bias.data.add_(0.5)
    m.repack()
    assert m.native_model(dev) is not h0


def test_exact_math_attribution(dev, golden):
    """where do the dB between the GPU path and the CPU oracle / reference goldens go?  Runs the chain with
    every approximation swapped for the exact operation, one switch at a time (plan flags B2D_PLAN_EXACT_*, conv_mode fp32,
    B2D_CONV_EXACT_GATES), and all together.  The table is printed (and written to gpurun_out/ when that directory exists);
    the asserts keep the default path at the agreed bar and make sure no switch makes things worse by more than noise."""
    import audio_denoising_b200 as adb
    from audio_denoising_b200 import _runtime as rt

    dsp, metrics, model, pipeline, synth = _oracle()
    m, sd, cfg = _our_model("good", dev)
    cases = []
    # the smoke() case: 2 clips of 1 s, oracle run here with the same injected initial phase
    noisy, _ = synth.make_batch(2, 16000, 16000)
    T = 1 + 16000 // 512
    init = synth.gl_init_angles((2, 513, T), seed=7)
    ref = pipeline.denoise_batch(noisy, model.GRUUNet2Oracle(sd, cfg), 1024, 512, 64, 16000, 32, 0.99, init)["wave"]
    cases.append(("smoke", 1024, 512, 16000, noisy, init, ref, True))
    c = golden("dsp_chain.npz")
    for tag in ("a", "b", "c"):
        n_fft, hop, sr, L = [int(v) for v in c[f"{tag}_cfg"]]
        cases.append((f"golden_{tag}", n_fft, hop, sr, torch.from_numpy(c[f"{tag}_noisy"]), torch.from_numpy(c[f"{tag}_init"]),
                      torch.from_numpy(c[f"{tag}_wave"]), False))
    switches = [("default", 0, "mma", False), ("exact_sqrt", rt.PLAN_EXACT_SQRT, "mma", False), ("exact_unit", rt.PLAN_EXACT_UNIT, "mma", False),
                ("exact_peak_div", rt.PLAN_EXACT_PEAK_DIV, "mma", False), ("fp32_invmel", rt.PLAN_FP32_INVMEL, "mma", False),
                ("fp32_conv", 0, "fp32", False), ("exact_gates", 0, "mma", True), ("all_exact", rt.PLAN_EXACT_ALL, "fp32", True),
                ("generic_kernels", rt.PLAN_GENERIC_KERNELS, "mma", False)]
    table = {}
    for name, n_fft, hop, sr, x, ini, want, norm in cases:
        row = {}
        for sw, flags, conv, gates in switches:
            m.conv_mode, m.exact_gates = conv, gates
            pipe = adb.DenoisePipeline(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=sr, n_iter=32, plan_flags=flags)
            wave, _ = pipe.denoise(x.to(dev), init_angles=ini.to(dev), normalise=norm)
            row[sw] = [round(float(v), 1) for v in metrics.si_sdr(wave.cpu(), want)]
        table[name] = row
    m.conv_mode, m.exact_gates = "mma", False
    print("\nSI-SDR(ours, oracle / reference golden) in dB per exact-math switch:")
    for name, row in table.items():
        for sw, v in row.items():
            print(f"  {name:10s} {sw:16s} {v}")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "exact_math_attribution.json"), "w") as f:
            json.dump(table, f, indent=1)
    for name, row in table.items():
        assert min(row["default"]) >= 40.0, (name, row["default"])
        assert min(row["all_exact"]) >= 40.0, (name, row["all_exact"])


# ---------------------------------------------------------------------------------------------- sibling models
@pytest.mark.parametrize("nm", [24, 27])
def test_momo3_matches_reference_golden(dev, golden, nm):
    """momo3.MOMO3 with the shipped MOMO3-4d4ea0 weights (hidden (16,16,16), paddings (1,0,1), delta-feature input) on the
    generic cell kernels against outputs of the reference module itself: full sequence, given hx, 2-D input, and chunked
    calls with hx / prev carried (momo3.py:272-324)."""
    import audio_denoising_b200 as adb

    _, metrics, *_ = _oracle()
    sd, cfg = load_weights("momo3")
    m = adb.MOMO3(**cfg)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    s = golden("siblings.npz")
    x = torch.from_numpy(s[f"momo_{nm}_x"]).to(dev)
    y, h = m(x)
    assert metrics.rel_l2(y.cpu(), torch.from_numpy(s[f"momo_{nm}_y"])) < 1e-5
    assert metrics.rel_l2(h.cpu(), torch.from_numpy(s[f"momo_{nm}_h"])) < 1e-5
    h0 = torch.from_numpy(s[f"momo_{nm}_h0"]).to(dev)
    y2, h2 = m(x, h0)
    assert metrics.rel_l2(y2.cpu(), torch.from_numpy(s[f"momo_{nm}_y_h0"])) < 1e-5
    assert metrics.rel_l2(h2.cpu(), torch.from_numpy(s[f"momo_{nm}_h_h0"])) < 1e-5
    assert torch.equal(h0.cpu(), torch.from_numpy(s[f"momo_{nm}_h0"]))  # the caller's hx is not mutated
    ya, ha = m(x[:, :4])
    yb, hb = m(x[:, 4:], ha, prev=x[:, 3:4])
    assert torch.equal(torch.cat([ya, yb], 1), y) and torch.equal(hb, h)  # chunked carry is bit-identical
    y2d, h2d = m(x[0])
    assert y2d.shape == (9, nm)
    assert metrics.rel_l2(y2d.cpu(), torch.from_numpy(s[f"momo_{nm}_y2d"])) < 1e-5
    with pytest.raises(_cabi_error()):
        m(torch.zeros(1, 3, 40, device=dev))  # 40 bins do not compress to the model's 3


def _cabi_error():
    from audio_denoising_b200 import _cabi

    return _cabi.B2DError


def test_gruunet_generic_configuration_matches_reference_golden(dev, golden):
    """gruunet.GRUUNet (gruunet.py: the same cell as gruunet2) in a configuration the tuned kernels do not cover -- hidden
    (8, 12), kernels (3, 5), paddings (1, 2), 5 bins, 4 Gaussians -- against the reference module's own output."""
    import audio_denoising_b200 as adb

    _, metrics, *_ = _oracle()
    s = golden("siblings.npz")
    cfg = json.loads(bytes(s["gru_cfg"]).decode())
    m = adb.GRUUNet(**cfg)
    assert not m.uses_tuned_kernels()
    sd = {k[len("gru_sd__"):]: torch.from_numpy(s[k]) for k in s.files if k.startswith("gru_sd__")}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    y, h = m(torch.from_numpy(s["gru_x"]).to(dev))
    assert metrics.rel_l2(y.cpu(), torch.from_numpy(s["gru_y"])) < 1e-5
    assert metrics.rel_l2(h.cpu(), torch.from_numpy(s["gru_h"])) < 1e-5
    with pytest.raises(NotImplementedError):
        adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=20, sample_rate=16000)


@pytest.mark.parametrize("name", CHECKPOINTS)
def test_gruunet_class_loads_gruunet2_checkpoints(dev, golden, name):
    """A checkpoint written for GRUUNet2 loads through the GRUUNet class (identical state_dict layout, gruunet.py:246-300) and
    reproduces the reference outputs on the tuned kernels; on the generic cell kernels the same weights agree to 1e-5."""
    import audio_denoising_b200 as adb
    from audio_denoising_b200._cell import ARCH_GRUUNET2, CellRunner

    _, metrics, *_ = _oracle()
    sd, cfg = load_weights(name)
    m = adb.GRUUNet(**cfg)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    io = golden("model_io.npz")
    x = torch.from_numpy(io["x"]).to(dev)
    y, h = m(x)
    assert metrics.rel_l2(y.cpu(), torch.from_numpy(io[f"{name}_y"])) < 1e-5
    assert metrics.rel_l2(h.cpu(), torch.from_numpy(io[f"{name}_h"])) < 1e-5
    hg = torch.zeros_like(h)
    yg = CellRunner(m, ARCH_GRUUNET2).forward(x, hg)
    assert metrics.rel_l2(yg.cpu(), torch.from_numpy(io[f"{name}_y"])) < 1e-5
    assert metrics.rel_l2(hg.cpu(), torch.from_numpy(io[f"{name}_h"])) < 1e-5


# ---------------------------------------------------------------------------------------------- fp32 backward
def test_backward_matches_torch_autograd_of_the_oracle(dev):
    """train() mode + autograd: gradients of a scalar loss w.r.t. the input, the initial hidden state and EVERY parameter (Gaussian
    channel weights and biases included) from the CUDA backward kernels (csrc/cell.cu) against torch autograd through the CPU
    oracle's per-frame cell (oracle/model.py = gruunet2.py:127-244), shipped configuration and shipped weights."""
    import audio_denoising_b200 as adb

    _, metrics, model, *_ = _oracle()
    sd, cfg = load_weights("good")
    m = adb.GRUUNet2(**cfg)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(3, 6, 64, generator=g) * 2).float()
    h0 = (torch.randn(3, 17, 4, generator=g) * 0.3).float()
    wy = torch.randn(3, 6, 64, generator=g)
    wh = torch.randn(3, 17, 4, generator=g)
    # ---- reference gradients: torch autograd through the oracle cell on the CPU ----
    orc = model.GRUUNet2Oracle(sd, cfg)
    names = [k for k in sd if not k.endswith("gs.offset")]
    for k in names:
        orc.sd[k] = orc.sd[k].clone().requires_grad_(True)
    xr, hr = x.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    h, outs = hr, []
    for t in range(x.shape[1]):
        o, h = orc.cell(xr[:, t, :].unsqueeze(1), h)
        outs.append(o)
    loss_ref = (torch.stack(outs, 1) * wy).sum() + (h * wh).sum()
    loss_ref.backward()
    # ---- ours ----
    xd, hd = x.to(dev).requires_grad_(True), h0.to(dev).requires_grad_(True)
    y, hT = m(xd, hd)
    assert y.requires_grad and hT.requires_grad
    loss = (y * wy.to(dev)).sum() + (hT * wh.to(dev)).sum()
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref)) + 1e-4
    loss.backward()
    assert metrics.rel_l2(xd.grad.cpu(), xr.grad) < 2e-5
    assert metrics.rel_l2(hd.grad.cpu(), hr.grad) < 2e-5
    params = dict(m.named_parameters())
    assert sorted(params) == sorted(names)
    for k in names:
        assert params[k].grad is not None, k
        assert metrics.rel_l2(params[k].grad.cpu(), orc.sd[k].grad) < 5e-5, (k, metrics.rel_l2(params[k].grad.cpu(), orc.sd[k].grad))
    # one optimiser step through the reference's TrainingContext wiring (server.py:86-99: AdamW over inner.parameters())
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    before = float(loss)
    for _ in range(5):
        opt.zero_grad()
        y, hT = m(xd.detach(), hd.detach())
        l2 = ((y - 0.5) ** 2).mean()
        l2.backward()
        opt.step()
    y2, _ = m(xd.detach(), hd.detach())
    assert float(((y2 - 0.5) ** 2).mean()) < float(l2) + 1e-6 and before == before
    # eval() / no_grad keeps using the fused inference kernels and returns tensors without a graph
    m.eval()
    ye, _ = m(xd.detach())
    assert not ye.requires_grad


@pytest.mark.parametrize("tag", ["momo", "g2", "gru"])
def test_backward_matches_reference_autograd_golden(dev, golden, tag):
    """Gradients from the CUDA backward kernels against torch autograd THROUGH THE REFERENCE MODULES THEMSELVES (tests/golden/
    make_golden_grads.py: momo3.MOMO3 with the shipped weights -- whose delta feature detaches the previous frame, momo3.py:278,287,
    so its input gradient is not the total derivative of the forward --, gruunet2.GRUUNet2 with GRUUNet2-good, gruunet.GRUUNet in
    the non-shipped configuration): loss value, d/dx, d/dh0 and d/d(every parameter)."""
    import audio_denoising_b200 as adb

    _, metrics, *_ = _oracle()
    gr = golden("grads.npz")
    if tag == "momo":
        sd, cfg = load_weights("momo3")
        m = adb.MOMO3(**cfg)
    elif tag == "g2":
        sd, cfg = load_weights("good")
        m = adb.GRUUNet2(**cfg)
    else:
        s = golden("siblings.npz")
        cfg = json.loads(bytes(s["gru_cfg"]).decode())
        sd = {k[len("gru_sd__"):]: torch.from_numpy(s[k]) for k in s.files if k.startswith("gru_sd__")}
        m = adb.GRUUNet(**cfg)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    x = torch.from_numpy(gr[f"{tag}_x"]).to(dev).requires_grad_(True)
    h0 = torch.from_numpy(gr[f"{tag}_h0"]).to(dev).requires_grad_(True)
    y, hT = m(x, h0)
    loss = (y * torch.from_numpy(gr[f"{tag}_wy"]).to(dev)).sum() + (hT * torch.from_numpy(gr[f"{tag}_wh"]).to(dev)).sum()
    ref_loss = float(gr[f"{tag}_loss"])
    assert abs(float(loss) - ref_loss) <= 1e-4 * abs(ref_loss) + 1e-4
    loss.backward()
    assert metrics.rel_l2(x.grad.cpu(), torch.from_numpy(gr[f"{tag}_gx"])) < 2e-5
    assert metrics.rel_l2(h0.grad.cpu(), torch.from_numpy(gr[f"{tag}_gh0"])) < 2e-5
    seen = 0
    for k, p in m.named_parameters():
        ref = torch.from_numpy(gr[f"{tag}_gp__{k}"])
        assert p.grad is not None, k
        if float(ref.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert metrics.rel_l2(p.grad.cpu(), ref) < 5e-5, (k, metrics.rel_l2(p.grad.cpu(), ref))
        seen += 1
    assert seen == sum(1 for f in gr.files if f.startswith(f"{tag}_gp__"))


def test_momo3_parameter_gradient_matches_finite_differences(dev):
    """Size-independent property: a directional derivative of the CUDA backward w.r.t. a weight tensor against central finite
    differences of the CUDA forward (fp32: 2e-2 relative).  (Not for the input: MOMO3 detaches the previous frame of the delta
    feature, so d loss / d x is not the total derivative; the golden test above pins it.)"""
    import audio_denoising_b200 as adb

    sd, cfg = load_weights("momo3")
    m = adb.MOMO3(**cfg)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(2, 5, 24, generator=g) * 2).to(dev)
    w = torch.randn(2, 5, 24, generator=g).to(dev)
    y, hT = m(x)
    (y * w).sum().backward()
    name, p = next((k, v) for k, v in m.named_parameters() if k.endswith("downs.1.conv.weight"))
    d = torch.randn(p.shape, generator=g).to(dev)
    d = d / d.norm()
    eps = 1e-2
    with torch.no_grad():
        base = p.detach().clone()

        def f(v):
            p.copy_(v)
            r = float((m(x)[0] * w).sum())
            p.copy_(base)
            return r

        fd = (f(base + eps * d) - f(base - eps * d)) / (2 * eps)
    an = float((p.grad * d).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(fd), abs(an)) + 1e-3, (name, fd, an)
