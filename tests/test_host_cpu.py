"""CPU: host-side logic that needs no GPU -- C-ABI export surface, drop-in module structure, argument
checking, sharding arithmetic and a 2-rank gloo run of the multi-process plumbing."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import CHECKPOINTS, ROOT, load_weights


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200denoise.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2d_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from audio_denoising_b200 import _build, _cabi

    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200denoise.h but not exported"
    # the ctypes table mirrors the header one to one
    assert sorted(_cabi.SIGNATURES) == declared
    assert _cabi.lib().b2d_version() >= 100


def test_c_abi_argument_errors_need_no_gpu():
    from audio_denoising_b200 import _cabi

    lib = _cabi.lib()
    out = ctypes.c_void_p()
    fb = (ctypes.c_float * 4)()
    assert lib.b2d_plan_create(1022, 511, 64, fb, fb, ctypes.byref(out)) == _cabi.ERR_UNSUPPORTED
    assert b"n_fft" in lib.b2d_last_error_string()
    assert lib.b2d_plan_create(1024, 0, 64, fb, fb, ctypes.byref(out)) == _cabi.ERR_BAD_ARG
    assert lib.b2d_plan_create(1024, 512, 64, None, fb, ctypes.byref(out)) == _cabi.ERR_BAD_ARG
    assert lib.b2d_peak(None, 1, 10, None, None) == _cabi.ERR_BAD_ARG
    assert lib.b2d_griffinlim_workspace_bytes(None, 1, 10) == 0
    with pytest.raises(_cabi.B2DError):
        _cabi.check(_cabi.ERR_BAD_ARG)


@pytest.mark.parametrize("name", CHECKPOINTS)
def test_gruunet2_dropin_state_dict_and_interface(name):
    import audio_denoising_b200 as adb

    sd, cfg = load_weights(name)
    m = adb.GRUUNet2(**cfg)
    assert list(m.state_dict().keys()) == list(sd.keys())  # same names, same order (optimizer.load_state_dict relies on it)
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(v.shape) for v in sd.values()]
    m.load_state_dict(sd)
    assert sum(v.numel() for v in m.state_dict().values()) == 15337  # 15 319 weights + 3 x 6 gs.offset buffer entries
    assert len(list(m.parameters())) == 18
    assert m.latent_size == 17 and m.num_compressed_bins == 4 and m.n_mels == 64
    assert m.get_config() == m.hparams and set(m.hparams) == set(cfg)
    m2 = adb.GRUUNet2.from_config(m.get_config())
    assert isinstance(m2, adb.GRUUNet2)
    # AdamW over parameters() can be built and its state loaded like TrainingContext.load does (server.py:90-91)
    opt = torch.optim.AdamW(m.parameters())
    opt.load_state_dict(opt.state_dict())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64))
    with pytest.raises(AssertionError):
        adb.GRUUNet2(4, 2, (17,) * 4, (3,) * 4, (2,) * 4, (1,) * 4)


def test_seeded_construction_matches_reference_layer_shapes():
    """Parameter holders are real Conv1d / ConvTranspose1d modules created in the reference's order."""
    import audio_denoising_b200 as adb
    from oracle import model as omodel

    torch.manual_seed(0)
    m = adb.GRUUNet2(**omodel.default_config())
    ref = omodel.random_state_dict()
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.items()}
    assert torch.equal(m.cell.input_gate.gs.offset, torch.linspace(0, 1, 6))


def test_transform_constructors_validate_like_torchaudio():
    import audio_denoising_b200 as adb

    adb.Spectrogram(power=None, n_fft=1536, win_length=1536, hop_length=768, window_fn=torch.hann_window)
    adb.MelScale(n_mels=64, n_stft=769, sample_rate=48000)
    adb.InverseMelScale(n_mels=64, n_stft=769, sample_rate=48000)
    adb.GriffinLim(n_fft=1536, win_length=1536, hop_length=768, window_fn=torch.hann_window, power=1.0)
    adb.InverseSpectrogram(n_fft=1024, win_length=1024, hop_length=512)
    with pytest.raises(ValueError):
        adb.GriffinLim(n_fft=1024, momentum=1.5)
    with pytest.raises(ValueError):
        adb.InverseMelScale(n_stft=513, driver="qr")
    with pytest.raises(NotImplementedError):
        adb.Spectrogram(n_fft=1024)  # power=2.0 default: not on the reference's path
    with pytest.raises(NotImplementedError):
        adb.Spectrogram(power=None, n_fft=1024, window_fn=torch.hamming_window)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        adb.Spectrogram(power=None, n_fft=1024)(torch.zeros(1, 4000))


def test_host_filterbank_is_bit_identical_to_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    from audio_denoising_b200._runtime import melscale_fbanks_htk

    for n_freqs, sr in [(513, 16000), (769, 48000), (321, 16000), (257, 16000)]:
        want = torchaudio.functional.melscale_fbanks(n_freqs, 0.0, float(sr // 2), 64, sr)
        assert torch.equal(melscale_fbanks_htk(n_freqs, 64, sr), want)


def test_utils_names_and_values():
    from audio_denoising_b200 import utils as u

    assert u.SR == 48000 and u.STDS.shape == (241,)
    x = torch.randn(2, 241, 5)
    assert torch.allclose(u.denormalize(u.normalize(x)), x, atol=1e-6)
    y = torch.randn(4, 7) * 3
    assert torch.allclose(u.unclamp(u.clamp(y)), y, atol=1e-4)
    z = torch.randn(3, 6, 5, dtype=torch.complex64)
    assert torch.equal(u.wrap_complex(u.unwrap_complex(z)), z)


def test_shard_ranges_cover_all_clips_once():
    from audio_denoising_b200.sharding import shard_range

    for n, w in [(10000, 8), (7, 4), (256, 1), (3, 8), (1250, 2)]:
        seen = []
        for r in range(w):
            lo, hi = shard_range(n, w, r)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi))
        assert seen == list(range(n))
        sizes = [shard_range(n, w, r)[1] - shard_range(n, w, r)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharded_driver():
    """world_size 2 on CPU (gloo): each rank takes its contiguous block of clips, no data-path collective,
    rank 0 gathers per-rank counters and the max over ranks of the step time."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", script]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK clips=11 max_ms=" in r.stdout


def test_host_fft_and_fast_path_index_math():
    """csrc/fft.cuh, csrc/gl_fast.cuh and csrc/gl_reg.cuh are __host__ __device__: run their unit tests on the CPU."""
    nvcc = "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    for name in ["fft_host_test", "gl_fast_host_test", "gl_reg_host_test"]:
        exe = os.path.join("/tmp", f"b2d_{name}")
        r = subprocess.run([nvcc, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", name + ".cu")], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout


def test_checkpoint_helpers_round_trip(tmp_path):
    """save_model / TrainingContext.load / loader heuristics (server.py:36-142, app3.py:46-119) against the B200 module."""
    import audio_denoising_b200 as adb

    sd, cfg = load_weights("good")
    ctx = adb.TrainingContext(adb.GRUUNet2, device="cpu", **cfg)
    ctx.inner.load_state_dict(sd)
    ctx.total_iters, ctx.batch_size = 123, 32
    ctx.test_loss_record = {0: 0.5, 1: 0.25}
    path = ctx.save(prefix=str(tmp_path))
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"last_epoch", "loss_record", "loss_metric", "last_target_name", "total_training_iters", "arch",
                       "last_batch_size", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "config"}
    assert ck["arch"] == "GRUUNet2" and ck["total_training_iters"] == 123
    name = os.path.basename(os.path.dirname(path))
    back = adb.TrainingContext.load(name, adb.GRUUNet2, prefix=str(tmp_path), device="cpu")
    assert back.total_iters == 123 and back.batch_size == 32 and back.best_eval_loss == 0.25
    for k, v in back.inner.state_dict().items():
        assert torch.equal(v, sd[k])
    model, dev = adb.load_denoising_model(path, device="cpu")
    assert isinstance(model, adb.GRUUNet2) and not model.training and dev.type == "cpu"
    assert adb.load_denoising_model(os.path.join(str(tmp_path), "missing.pth")) == (None, None)  # swallowed like app3.py:118-119
    with pytest.raises(FileNotFoundError):
        adb.load_denoising_model(os.path.join(str(tmp_path), "missing.pth"), strict_errors=True)
    bare = os.path.join(str(tmp_path), "bare.pth")
    torch.save(dict(sd), bare)  # bare state dict + fallback config (app3.py:71-88)
    model2, _ = adb.load_denoising_model(bare, fallback_config=cfg, device="cpu")
    assert model2 is not None
    assert issubclass(adb.GRUUNet, adb.GRUUNet2)


@pytest.mark.parametrize("orig,new", [(44100, 48000), (48000, 44100), (16000, 48000), (48000, 16000), (22050, 16000)])
def test_resample_table_bit_identical_to_torchaudio(orig, new):
    """Host-built sinc-Hann polyphase table == torchaudio's (TA:functional/functional.py:1340-1402)."""
    import math
    import torchaudio
    from audio_denoising_b200.transforms import _sinc_hann_table

    g = math.gcd(orig, new)
    table, width = _sinc_hann_table(orig // g, new // g, 6, 0.99)
    ref, ref_width = torchaudio.functional.functional._get_sinc_resample_kernel(orig, new, g)
    assert width == ref_width and torch.equal(table, ref[:, 0, :])


def test_resample_and_ingest_refuse_cpu_tensors():
    import audio_denoising_b200 as adb

    with pytest.raises((TypeError, ValueError, RuntimeError)):
        adb.Resample(44100, 48000)(torch.zeros(100))
    with pytest.raises(TypeError):
        adb.pcm16_to_float(torch.zeros(10, dtype=torch.int16))
    with pytest.raises(NotImplementedError):
        adb.Resample(44100, 48000, resampling_method="sinc_interp_kaiser")


def test_momo3_dropin_state_dict_and_interface():
    """momo3.MOMO3 drop-in: constructor, state_dict keys / shapes / order of the shipped MOMO3-4d4ea0 checkpoint, hparams helpers;
    CPU tensors are refused (no fallback)."""
    import audio_denoising_b200 as adb
    from conftest import load_weights

    sd, cfg = load_weights("momo3")
    m = adb.MOMO3(**cfg)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(v.shape) for v in sd.values()]
    m.load_state_dict(sd)
    assert m.get_config() == m.hparams and adb.MOMO3.from_config(m.get_config()).latent_size == 16
    assert m.num_compressed_bins == 3
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 24))
    g = adb.GRUUNet2(4, 1, (8, 12), (3, 5), (2, 2), (1, 2))
    assert not g.uses_tuned_kernels() and adb.GRUUNet2(4, 1, (17,) * 4, (3,) * 4, (2,) * 4, (1,) * 4).uses_tuned_kernels()


def test_derived_seed_stream_is_nonzero_62_bit_and_decorrelated():
    """_runtime.derive_seed: the per-hop seeds of the streaming denoiser (one torch draw per denoiser, splitmix64 per hop)."""
    from audio_denoising_b200._runtime import derive_seed

    seeds = [derive_seed(0x1234_5678_9ABC, k) for k in range(4096)]
    assert len(set(seeds)) == len(seeds) and all(0 < s < 2 ** 62 for s in seeds)
    assert derive_seed(0x1234_5678_9ABC, 7) == seeds[7]  # a pure function of (base, index)
    assert derive_seed(0x1234_5678_9ABD, 7) != seeds[7]
    # consecutive seeds differ in about half of their bits (no counter-like structure reaches the in-kernel hash)
    flips = [bin(a ^ b).count("1") for a, b in zip(seeds, seeds[1:])]
    assert 26 < sum(flips) / len(flips) < 36
    assert derive_seed(0, 0) != 0
