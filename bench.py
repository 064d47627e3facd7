#!/usr/bin/env python
"""bench.py -- headline benchmark of the belacks/audio-denoising hot path on B200.

Metric (BASELINE.json): audio-seconds denoised per wall-second for
``STFT -> Mel -> GRUUNet2 -> inverse Mel -> Griffin-Lim(32) -> iSTFT`` on BASELINE ``configs[1]``:
a batch of 256 x 4 s 16 kHz clips per GPU (n_fft 1024, hop 512, 64 mel), synthetic noisy clips,
shipped ``GRUUNet2-good`` weights (tests/golden/weights_good.npz); plus p50 / p99 per-hop latency of the
streaming mode (configs[2]) and the sharded 10 s corpus (configs[3]) as sub-records of the same line.

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [...]                     the reference's CPU path (oracle port: torchaudio + restated
                                                             model) on the box's host cores, bounded sample per step

One "step" = one pass of the whole chain over one batch.  ``value`` is measured with the inputs resident in
HBM (CUDA events, max over ranks); ``e2e`` goes through the public host API (``DenoisePipeline.denoise_host``) with
pinned host buffers, H2D and D2H inside the timed region, on the link format the reference exchanges: int16 PCM in,
int16 PCM out (app3.py:168-172, :244-245) -- the conversions are part of the timed work in both arms; the float32
link is reported beside it (``e2e.float32_io``).  The working set of a step (~0.5 GB of Griffin-Lim state per batch)
is far larger than the 126 MB L2, so no explicit L2 flush is needed between iterations (stated in ``config.l2``).
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
STREAM_GEOMETRIES = [(16000, 640, 320), (48000, 1536, 768)]  # configs[2]: 20 ms hop @ 16 kHz; the reference-native app3.py:29-33
CORPUS_CLIPS, CORPUS_SECONDS, CORPUS_BATCH = 10000, 10, 128  # configs[3]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=256, help="clips per CPU-baseline pass (256 = the whole batch, ~4 s of host time per pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the streaming / corpus sub-records")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 40)")
    ap.add_argument("--stream-hops", type=int, default=1000)
    ap.add_argument("--corpus-clips", type=int, default=CORPUS_CLIPS)
    return ap.parse_args()


def config_dict(batch: int, world: int) -> dict:
    """The workload description both arms print (identical dict, so the driver's same_config check can hold)."""
    return {"workload": WORKLOAD, "clips_per_gpu": batch, "clip_seconds": SECONDS, "sr": SR, "n_fft": N_FFT, "hop": HOP,
            "n_mels": N_MELS, "gl_iters": N_ITER, "weights": "GRUUNet2-good (shipped checkpoint)",
            "link": "int16 PCM host buffers in and out (app3.py:168-172, :244-245); value = device-resident float32",
            "parallelism": f"dp{world} (independent clips, no collective on the data path)",
            "l2": "per-step working set ~0.5 GB >> 126 MB L2; no explicit flush"}


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


def to_pcm16(x):
    """(clip(x, -1, 1) * 32767).astype(int16), app3.py:244-245."""
    import torch

    return (x.clamp(-1.0, 1.0) * 32767.0).to(torch.int16)


def load_model_weights(name: str = "good"):
    import numpy as np
    import torch

    z = np.load(os.path.join(ROOT, "tests", "golden", f"weights_{name}.npz"))
    cfg = json.loads(bytes(z["__config__"]).decode())
    sd = {k: torch.from_numpy(z[k].copy()) for k in z.files if k != "__config__"}
    return sd, cfg


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: oracle (torchaudio transforms' arithmetic + restated GRUUNet2) on host cores.
# The only place (with tests/ and smoke()) that executes oracle/ -- as the reference's CPU path, never as the product.
# ------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self):
        from oracle import model as omodel

        sd, cfg = load_model_weights()
        self.orc = omodel.GRUUNet2Oracle(sd, cfg)

    def pass_pcm16(self, pcm):
        """One pass of the reference path over an int16 batch: int16 -> float (app3.py:172), chain, float -> int16 (:244-245)."""
        import torch

        from oracle import pipeline as opipe

        x = pcm.to(torch.float32) / 32767.0
        wave = opipe.denoise_batch(x, self.orc, N_FFT, HOP, N_MELS, SR, N_ITER, 0.99, None)["wave"]
        return to_pcm16(wave)

    def seconds(self, pcm, reps: int, threads: int) -> float:
        import torch

        torch.set_num_threads(threads)
        best = float("inf")
        for i in range(reps + 1):  # first pass is the warm-up
            t0 = time.perf_counter()
            self.pass_pcm16(pcm)
            dt = time.perf_counter() - t0
            if i > 0:
                best = min(best, dt)
        return best


def cpu_streaming_p50(sr: int, n_fft: int, hop: int, hops: int = 24) -> float:
    """p50 per-hop latency (ms) of the reference's streaming loop (app3.py:167-226) restated in oracle.StreamingOracle."""
    import numpy as np

    from oracle import model as omodel
    from oracle import pipeline as opipe

    sd, cfg = load_model_weights("dari_tult2")
    so = opipe.StreamingOracle(omodel.GRUUNet2Oracle(sd, cfg), n_fft=n_fft, hop=hop, n_mels=N_MELS, sample_rate=sr, n_iter=N_ITER)
    rng = np.random.default_rng(0)
    so.push((rng.standard_normal(n_fft) * 0.1).astype(np.float32))  # fills the window: first hop = warm-up
    lat = []
    for _ in range(hops):
        chunk = (rng.standard_normal(hop) * 0.1).astype(np.float32)
        t0 = time.perf_counter()
        so.push(chunk)
        lat.append((time.perf_counter() - t0) * 1e3)
    return float(statistics.median(lat))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    arm = CpuArm()
    n = args.cpu_sample
    pcm = to_pcm16(synth_batch(n, SR * SECONDS))
    steps = max(1, min(args.steps, 4))  # each step = one pass over the sample (~4 s for 256 clips on 16 cores)
    warm = max(1, min(args.warmup, 1))
    for _ in range(warm):
        arm.pass_pcm16(pcm)
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.pass_pcm16(pcm)
    dt = (time.perf_counter() - t0) / steps
    value = n * SECONDS / dt
    sample = (f"{n} clips x {SECONDS} s per step (of the 256-clip batch), {steps} timed steps, int16 PCM in -> float -> chain -> int16 PCM out, "
              "oracle port (torchaudio arithmetic + per-frame restated GRUUNet2) on torch CPU fp32")
    line = {
        "impl": "reference", "metric": "audio-sec/sec", "value": round(value, 2), "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.batch, world),
        "cpu_baseline": {"value": round(value, 2), "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model_name()},
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


def measure_streaming(adb, dev, hops: int, with_cpu: bool):
    """configs[2]: one session, per-hop latency from the chunk in pinned host memory to the emitted hop back in pinned
    host memory (H2D + whole chain with carried GRU state + D2H + sync: what app3.py:189-215 spends per hop)."""
    import numpy as np
    import torch

    sd, cfg = load_model_weights("dari_tult2")  # the checkpoint app3.py loads
    m = adb.GRUUNet2(**cfg)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    rows = []
    warm = 100
    for sr, n_fft, hop in STREAM_GEOMETRIES:
        sdn = adb.StreamingDenoiser(m, n_fft=n_fft, hop_length=hop, n_mels=N_MELS, sample_rate=sr, n_iter=N_ITER, sessions=1, device=dev)
        rng = np.random.default_rng(0)
        sig = (rng.standard_normal((1, n_fft + hop * (hops + warm))) * 0.1).astype(np.float32)
        lat = []
        for i in range(hops + warm):
            win = sig[:, i * hop: i * hop + n_fft]
            t0 = time.perf_counter()
            sdn.step(win)
            dt = time.perf_counter() - t0
            if i >= warm:
                lat.append(dt * 1e3)
        lat = np.array(lat)
        row = {"geometry": f"{sr} Hz, n_fft {n_fft}, hop {hop} ({1000.0 * hop / sr:.0f} ms)", "sessions": 1, "hops": hops,
               "p50_ms": round(float(np.percentile(lat, 50)), 4), "p99_ms": round(float(np.percentile(lat, 99)), 4),
               "includes": "H2D of the chunk, STFT..Griffin-Lim(32)..overlap-add with carried hx, D2H of the hop, stream sync"}
        if with_cpu:
            row["cpu_p50_ms"] = round(cpu_streaming_p50(sr, n_fft, hop), 2)
            row["cpu_kind"] = "port (oracle.StreamingOracle = app3.py recv loop on torch CPU)"
        rows.append(row)
        del sdn
    return rows


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

    def max_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms_max = max_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * B * SECONDS / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel: the fused Griffin-Lim iteration ---------------------------
    # average launch duration measured live with CUDA events on the launching stream:
    # (time of GL with n_iter=32) - (time with n_iter=0: init iSTFT + stitch only), divided by 32 launches.
    roof = None
    if rank == 0:
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
                j = json.load(open(tr))
                roof["traffic"] = j.get("dram_bytes_per_launch")
                roof["traffic_source"] = "committed ncu --set full capture, not measured in this run: " + str(j.get("source"))
                if j.get("kernel_reads_bytes_per_launch"):
                    roof["kernel_bytes_per_launch"] = j["kernel_reads_bytes_per_launch"]
            except Exception:
                pass

    # ---- end to end through the public host API (pinned host in, pinned host out) ---------------------
    # Every step uploads its own input batch from pinned host memory and downloads its own result; consecutive steps
    # are software-pipelined (upload of step i+1 / download of step i-1 overlap the kernels of step i), as a corpus
    # driver would run it.  The clock stops when the last result is in host memory.  Headline link = int16 PCM
    # (what the reference's recv exchanges); the float32 link is measured the same way beside it.
    e2e_steps = args.e2e_steps or max(4, min(args.steps, 40))

    def e2e_run(host_in, host_out):
        for i in range(3):
            pipe.denoise_host(host_in[i % 2], host_out[i % 2], wait=False)
        pipe.host_synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            pipe.denoise_host(host_in[i % 2], host_out[i % 2], wait=False)
        pipe.host_synchronize()
        return world * B * SECONDS / max_ranks((time.perf_counter() - t0) / e2e_steps)

    pcm_in = [to_pcm16(noisy_host).pin_memory(), to_pcm16(synth_batch(B, L, seed=4321 + rank)).pin_memory()]
    pcm_out = [torch.empty((B, Lout), dtype=torch.int16, pin_memory=True) for _ in range(2)]
    e2e_pcm16 = e2e_run(pcm_in, pcm_out)
    f32_in = [noisy_host, noisy_host.clone().pin_memory()]
    f32_out = [torch.empty((B, Lout), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    e2e_f32 = e2e_run(f32_in, f32_out)
    e2e_pcm16 = max(e2e_pcm16, e2e_run(pcm_in, pcm_out))  # second pass after the float32 one: report the better of two

    # host link bandwidth seen by this process (explains e2e when the PCIe link, not the kernels, is the limit)
    def copy_gbs(dst, src):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize(dev)
        return round(3 * src.numel() * src.element_size() / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
    link = {"h2d_gbs": copy_gbs(noisy, noisy_host), "d2h_gbs": copy_gbs(f32_out[0], wave)}
    del f32_in, f32_out

    # ---- configs[3]: the sharded corpus, int16 link, every clip through pinned host memory -------------
    corpus = None
    if not args.no_extras:
        from audio_denoising_b200.sharding import shard_range

        Lc = SR * CORPUS_SECONDS
        lo, hi = shard_range(args.corpus_clips, world, rank)
        Bt = CORPUS_BATCH
        block = to_pcm16(synth_batch(Bt, Lc, seed=77 + rank)).pin_memory()  # host synthesis excluded: one block re-used
        outs = [torch.empty((Bt, pipe.out_length(Lc)), dtype=torch.int16, pin_memory=True) for _ in range(2)]
        for i in range(2):
            pipe.denoise_host(block, outs[i % 2], wait=False)
        pipe.host_synchronize()
        barrier()
        t0 = time.perf_counter()
        done, i = 0, 0
        while done < hi - lo:
            nb = min(Bt, hi - lo - done)
            pipe.denoise_host(block[:nb], outs[i % 2][:nb], wait=False)
            done += nb
            i += 1
        pipe.host_synchronize()
        sec = max_ranks(time.perf_counter() - t0)
        corpus = {"clips": args.corpus_clips, "clip_seconds": CORPUS_SECONDS, "seconds": round(sec, 4),
                  "audio_s_per_s": round(args.corpus_clips * CORPUS_SECONDS / sec, 1), "batch": Bt,
                  "sharding": f"contiguous block of {hi - lo} clips per rank, {world} rank(s), no collective",
                  "includes": "int16 PCM pinned host -> device -> pinned host for every clip (host synthesis excluded)"}
        del block, outs

    # ---- configs[2]: streaming latency (rank 0; one session) and the CPU comparators ---------------------
    streaming = None
    if rank == 0 and not args.no_extras:
        streaming = measure_streaming(adb, dev, args.stream_hops, with_cpu=not args.no_cpu_baseline)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = args.cpu_sample
        arm = CpuArm()
        dt = arm.seconds(to_pcm16(synth_batch(n, L)), 2, threads)
        n1 = max(1, min(8, n))
        dt1 = arm.seconds(to_pcm16(synth_batch(n1, L)), 1, 1)
        cpu = {"value": round(n * SECONDS / dt, 2), "unit": "audio-s/s", "cores": threads, "kind": "port",
               "cpu_model": cpu_model_name(), "value_1thread": round(n1 * SECONDS / dt1, 2),
               "sample": f"{n} clips x {SECONDS} s (of the 256-clip batch), best of 2 passes after 1 warm-up (~{3 * dt:.0f} s of host time), "
                         f"int16 PCM in/out; 1-thread figure on {n1} clips; "
                         "oracle port (torchaudio arithmetic + per-frame restated GRUUNet2) on torch CPU fp32"}

    if rank == 0:
        line = {
            "metric": "audio-sec/sec", "value": round(value, 1), "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(B, world),
            "clocks": clocks,
            "e2e": {"value": round(e2e_pcm16, 1), "unit": "audio-s/s", "link": "int16 PCM", "h2d_bytes_per_step": B * L * 2,
                    "d2h_bytes_per_step": B * Lout * 2, "steps": e2e_steps, "pipelined": True, **link, "host_numa_binding": numa,
                    "float32_io": {"value": round(e2e_f32, 1), "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * Lout * 4}},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "streaming": streaming,
            "corpus": corpus,
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
