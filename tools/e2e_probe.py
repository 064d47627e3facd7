"""Where the host-to-host step goes at N ranks (torchrun): the denoise_host pipeline re-stated with timing events -- per step
the duration of the upload, of the kernels and of the download, and when each started relative to the step before -- for
several staging-ring depths.  int16 PCM link, BASELINE config 2 batch."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import audio_denoising_b200 as adb
import bench

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sd, cfg = bench.load_model_weights()
model = adb.GRUUNet2(**cfg); model.load_state_dict(sd); model = model.to(dev).eval()
B, L = 256, 64000
steps = 24


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(depth):
    pipe = adb.DenoisePipeline(model, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=32)
    pipe.host_ring_depth = depth
    Lout = pipe.out_length(L)
    hin = [torch.randint(-8000, 8000, (B, L), dtype=torch.int16).pin_memory() for _ in range(2)]
    hout = [torch.empty((B, Lout), dtype=torch.int16).pin_memory() for _ in range(depth)]
    for i in range(4):
        pipe.denoise_host(hin[i % 2], hout[i % depth], wait=False)
    pipe.host_synchronize(); barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        pipe.denoise_host(hin[i % 2], hout[i % depth], wait=False)
    pipe.host_synchronize()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timeline(depth):
    """the same loop by hand with timing events (one rank's view)"""
    pipe = adb.DenoisePipeline(model, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=32)
    Lout = pipe.out_length(L)
    hin = [torch.randint(-8000, 8000, (B, L), dtype=torch.int16).pin_memory() for _ in range(2)]
    hout = [torch.empty((B, Lout), dtype=torch.int16).pin_memory() for _ in range(depth)]
    xin = [torch.empty((B, L), dtype=torch.int16, device=dev) for _ in range(depth)]
    res = [torch.empty((B, Lout), dtype=torch.int16, device=dev) for _ in range(depth)]
    h2d, d2h, comp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    E = lambda: torch.cuda.Event(enable_timing=True)
    in_free, out_free = [None] * depth, [None] * depth
    marks = []
    pipe.denoise_pcm16(xin[0], out=res[0]); barrier()
    origin = E(); origin.record(comp)
    h2d.wait_event(origin); d2h.wait_event(origin)
    for i in range(steps):
        s = i % depth
        a0, a1, c0, c1, b0, b1 = E(), E(), E(), E(), E(), E()
        with torch.cuda.stream(h2d):
            if in_free[s] is not None: h2d.wait_event(in_free[s])
            a0.record(h2d); xin[s].copy_(hin[i % 2], non_blocking=True); a1.record(h2d)
        comp.wait_event(a1)
        if out_free[s] is not None: comp.wait_event(out_free[s])
        c0.record(comp); pipe.denoise_pcm16(xin[s], out=res[s]); c1.record(comp)
        in_free[s] = c1
        with torch.cuda.stream(d2h):
            d2h.wait_event(c1)
            b0.record(d2h); hout[s].copy_(res[s], non_blocking=True); b1.record(d2h)
        out_free[s] = b1
        marks.append((a0, a1, c0, c1, b0, b1))
    torch.cuda.synchronize()
    rows = [[round(origin.elapsed_time(e), 3) for e in m] for m in marks]
    return rows


out = {"n_gpus": world}
for depth in (2, 3, 4):
    ms = run(depth) * 1e3
    out[f"ring{depth}_ms_per_step"] = round(ms, 3)
    out[f"ring{depth}_audio_s_per_s"] = round(world * B * 4 / (ms * 1e-3))
rows = timeline(2)
if rank == 0:
    r = np.array(rows[8:])
    out["timeline_ring2_ms"] = dict(h2d=round(float((r[:, 1] - r[:, 0]).mean()), 3), compute=round(float((r[:, 3] - r[:, 2]).mean()), 3),
                                    d2h=round(float((r[:, 5] - r[:, 4]).mean()), 3), step=round(float(np.diff(r[:, 3]).mean()), 3),
                                    compute_gap=round(float((r[1:, 2] - r[:-1, 3]).mean()), 3),
                                    h2d_start_after_prev_compute_end=round(float((r[2:, 0] - r[:-2, 3]).mean()), 3))
    out["timeline_last_steps"] = rows[-3:]
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
