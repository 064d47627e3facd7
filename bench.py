#!/usr/bin/env python
"""bench.py -- headline benchmark of the belacks/audio-denoising hot path on B200.

Metric (BASELINE.json): audio-seconds denoised per wall-second for
``STFT -> Mel -> GRUUNet2 -> inverse Mel -> Griffin-Lim(32) -> iSTFT`` on BASELINE ``configs[1]``:
a batch of 256 x 4 s 16 kHz clips per GPU (n_fft 1024, hop 512, 64 mel), synthetic noisy clips,
shipped ``GRUUNet2-good`` weights (tests/golden/weights_good.npz).

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [...]                     the reference's CPU path (oracle port: torchaudio + restated
                                                             model) on the box's host cores, bounded sample per step

One "step" = one pass of the whole chain over one batch.  ``value`` is measured with the inputs resident in
HBM (CUDA events, max over ranks); ``e2e`` goes through the public host API with pinned host buffers, H2D and
D2H inside the timed region.  The working set of a step (~0.6 GB of spectrogram state per batch) is far larger
than the 126 MB L2, so no explicit L2 flush is needed between iterations (stated in ``config.l2``).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, SECONDS, N_FFT, HOP, N_MELS, N_ITER = 16000, 4, 1024, 512, 64, 32
WORKLOAD = "batch 256 x 4 s 16 kHz clips, GRUUNet2 + 32 Griffin-Lim iterations (BASELINE configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--conv-mode", default="mma", choices=["fp32", "tf32x3", "tf32", "mma", "mma_tf32"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="clips per CPU-baseline pass (256 = the whole batch, ~4 s of host time per pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    return ap.parse_args()


def synth_batch(n: int, length: int, seed: int = 1234):
    """Seeded synthetic noisy clips (tone + chirp + AM harmonic stack + white noise, peak-normalised), host float32.

    Kernel timing on this path is data independent; the generator only has to produce plausible audio."""
    import math

    import torch

    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length, dtype=torch.float32) / SR
    u = torch.rand(n, 5, generator=g)
    f0 = (100.0 + 1900.0 * u[:, 0:1]) * t
    f1, f2 = 100.0 + 400.0 * u[:, 1:2], 1000.0 + 5000.0 * u[:, 2:3]
    dur = length / SR
    x = 0.5 * torch.sin(2 * math.pi * f0) + 0.5 * torch.sin(2 * math.pi * (f1 + (f2 - f1) * t / (2 * dur)) * t)
    pitch = (120.0 + 30.0 * torch.sin(2 * math.pi * 3.0 * t)) * t
    voiced = sum(torch.sin(2 * math.pi * k * pitch) / k for k in range(1, 6))
    x = x + 0.3 * voiced * (0.5 + 0.5 * torch.sin(2 * math.pi * 4.0 * t))
    snr = 10.0 ** (-(20.0 * u[:, 3:4]) / 20.0)
    x = x + snr * x.std(dim=1, keepdim=True) * torch.randn(n, length, generator=g)
    return (x / x.abs().amax(dim=1, keepdim=True).clamp_min(1e-6)).contiguous()


def load_model_weights():
    import numpy as np
    import torch

    z = np.load(os.path.join(ROOT, "tests", "golden", "weights_good.npz"))
    cfg = json.loads(bytes(z["__config__"]).decode())
    sd = {k: torch.from_numpy(z[k].copy()) for k in z.files if k != "__config__"}
    return sd, cfg


# ------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: oracle (torchaudio transforms' arithmetic + restated GRUUNet2) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_pass_seconds(n_clips: int, reps: int, threads: int):
    import torch

    from oracle import model as omodel
    from oracle import pipeline as opipe

    torch.set_num_threads(threads)
    sd, cfg = load_model_weights()
    orc = omodel.GRUUNet2Oracle(sd, cfg)
    noisy = synth_batch(n_clips, SR * SECONDS)
    best = float("inf")
    for i in range(reps + 1):  # first pass is the warm-up
        t0 = time.perf_counter()
        opipe.denoise_batch(noisy, orc, N_FFT, HOP, N_MELS, SR, N_ITER, 0.99, None)
        dt = time.perf_counter() - t0
        if i > 0:
            best = min(best, dt)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import model as omodel
    from oracle import pipeline as opipe

    sd, cfg = load_model_weights()
    orc = omodel.GRUUNet2Oracle(sd, cfg)
    n = args.cpu_sample
    noisy = synth_batch(n, SR * SECONDS)
    steps = max(1, min(args.steps, 4))  # each step = one pass over the sample (~4 s for 256 clips on 16 cores)
    warm = max(1, min(args.warmup, 1))
    for _ in range(warm):
        opipe.denoise_batch(noisy, orc, N_FFT, HOP, N_MELS, SR, N_ITER, 0.99, None)
    t0 = time.perf_counter()
    for _ in range(steps):
        opipe.denoise_batch(noisy, orc, N_FFT, HOP, N_MELS, SR, N_ITER, 0.99, None)
    dt = (time.perf_counter() - t0) / steps
    value = n * SECONDS / dt
    sample = f"{n} clips x {SECONDS} s per step (of the 256-clip batch), {steps} timed steps, torch CPU fp32"
    line = {
        "impl": "reference", "metric": "audio-sec/sec", "value": round(value, 2), "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "n_fft": N_FFT, "hop": HOP, "n_mels": N_MELS, "gl_iters": N_ITER, "sr": SR},
        "cpu_baseline": {"value": round(value, 2), "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import audio_denoising_b200 as adb
    from audio_denoising_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from audio_denoising_b200.sharding import bind_host_to_gpu

    numa = bind_host_to_gpu(local)  # before any pinned allocation: staging buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    adb.native_library()

    B, L = args.batch, SR * SECONDS
    sd, cfg = load_model_weights()
    model = adb.GRUUNet2(**cfg)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    model.conv_mode = args.conv_mode
    pipe = adb.DenoisePipeline(model, n_fft=N_FFT, hop_length=HOP, n_mels=N_MELS, sample_rate=SR, n_iter=N_ITER)
    T, F = pipe.num_frames(L), N_FFT // 2 + 1
    Lout = pipe.out_length(L)

    noisy_host = synth_batch(B, L, seed=1234 + rank).pin_memory()
    noisy = noisy_host.to(dev)
    wave = torch.empty((B, Lout), dtype=torch.float32, device=dev)

    def step():
        pipe.denoise(noisy, out=wave)  # rand_init=True: fresh in-kernel random initial phase every step, like the reference

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _cabi.launch_count() - launches0
    clocks = sampler.finish()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * SECONDS / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel: the fused Griffin-Lim iteration ---------------------------
    # average launch duration measured live with CUDA events on the launching stream:
    # (time of GL with n_iter=32) - (time with n_iter=0: init iSTFT + stitch only), divided by 32 launches.
    roof = None
    if rank == 0:
        import ctypes as C

        lib = _cabi.lib()
        mag = torch.rand((B, T, pipe.plan.frame_stride), dtype=torch.float32, device=dev)
        nbytes = lib.b2d_griffinlim_workspace_bytes(pipe.plan.handle, B, T)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream

        def gl(n_iter):
            _cabi.check(lib.b2d_griffinlim_frames(pipe.plan.handle, mag.data_ptr(), None, 12345, B, T, n_iter, 0.99, None,
                                                  wave.data_ptr(), ws.data_ptr(), ws.numel(), st))

        def timed(n_iter, reps):
            for _ in range(2):
                gl(n_iter)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                gl(n_iter)
            b.record()
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / reps

        reps = max(5, min(args.steps, 20))
        per_launch_ms = (timed(N_ITER, reps) - timed(0, reps)) / N_ITER
        alg_bytes = B * (20 * F * T + 8 * Lout)  # per iteration per clip 20 F T + 8 L_out
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "griffin-lim fused iteration", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": which,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": round(per_launch_ms, 5)}
        tr = os.path.join(ROOT, "profiles", "gl_traffic.json")
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
            except Exception:
                pass

    # ---- end to end through the public host API (pinned host in, pinned host out) ---------------------
    # Every step uploads its own input batch from pinned host memory and downloads its own result; consecutive steps
    # are software-pipelined (upload of step i+1 / download of step i-1 overlap the kernels of step i), as a corpus
    # driver would run it.  The clock stops when the last result is in host memory.
    host_in = [noisy_host, noisy_host.clone().pin_memory()]
    host_out = [torch.empty((B, Lout), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    e2e_steps = args.e2e_steps or max(4, min(args.steps, 40))
    for i in range(3):
        pipe.denoise_host(host_in[i % 2], host_out[i % 2], wait=False)
    pipe.host_synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        pipe.denoise_host(host_in[i % 2], host_out[i % 2], wait=False)
    pipe.host_synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * SECONDS / float(t2.item())
    # the same with int16 PCM on the link (what app3.py's recv exchanges, app3.py:168-172 / 244-245): conversions run on the GPU
    pcm_in = [(h.clamp(-1, 1) * 32767).to(torch.int16).pin_memory() for h in host_in]
    pcm_out = [torch.empty((B, Lout), dtype=torch.int16, pin_memory=True) for _ in range(2)]
    for i in range(3):
        pipe.denoise_host(pcm_in[i % 2], pcm_out[i % 2], wait=False)
    pipe.host_synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        pipe.denoise_host(pcm_in[i % 2], pcm_out[i % 2], wait=False)
    pipe.host_synchronize()
    t3 = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    e2e_pcm16 = world * B * SECONDS / float(t3.item())
    if os.environ.get("B2D_BENCH_DEBUG"):
        for tag, (hi_, ho_) in {"fp32-again": (host_in, host_out), "pcm-again": (pcm_in, pcm_out)}.items():
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                pipe.denoise_host(hi_[i % 2], ho_[i % 2], wait=False)
            pipe.host_synchronize()
            print(f"[debug] {tag}: {(time.perf_counter() - t0) / e2e_steps * 1e3:.3f} ms/step; first fp32 {e2e_s * 1e3:.3f}", file=sys.stderr, flush=True)
    # host link bandwidth seen by this process (explains e2e when the PCIe link, not the kernels, is the limit)
    def copy_gbs(dst, src):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize(dev)
        return round(3 * src.numel() * 4 / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
    link = {"h2d_gbs": copy_gbs(noisy, noisy_host), "d2h_gbs": copy_gbs(host_out[0], wave)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = args.cpu_sample
        dt = cpu_pass_seconds(n, 2, threads)
        cpu = {"value": round(n * SECONDS / dt, 2), "unit": "audio-s/s", "cores": threads, "kind": "port",
               "sample": f"{n} clips x {SECONDS} s (of the 256-clip batch), best of 2 passes after 1 warm-up (~{3 * dt:.0f} s of host time), "
                         "oracle port (torchaudio arithmetic + per-frame restated GRUUNet2) on torch CPU fp32"}

    if rank == 0:
        line = {
            "metric": "audio-sec/sec", "value": round(value, 1), "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": B, "clip_seconds": SECONDS, "sr": SR, "n_fft": N_FFT, "hop": HOP,
                       "n_mels": N_MELS, "gl_iters": N_ITER, "conv_mode": args.conv_mode, "weights": "GRUUNet2-good (shipped checkpoint)",
                       "parallelism": f"dp{world} (independent clips, no collective on the data path)",
                       "l2": "per-step working set ~0.6 GB >> 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "audio-s/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * Lout * 4,
                    "steps": e2e_steps, "pipelined": True, **link, "host_numa_binding": numa,
                    "int16_pcm_io": {"value": round(e2e_pcm16, 1), "h2d_bytes_per_step": B * L * 2, "d2h_bytes_per_step": B * Lout * 2}},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
