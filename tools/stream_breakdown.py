"""Where a streaming hop's latency goes (BASELINE config 3): host-visible step() time, device time of the captured graph
(CUDA events around the replay), an empty-graph replay + sync (host floor), and each kernel alone (events, eager launches)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import audio_denoising_b200 as adb
from conftest import load_weights

hops = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda:0")
sd, cfg = load_weights("dari_tult2")
m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
for sr, n_fft, hop in [(16000, 640, 320), (48000, 1536, 768)]:
    sdn = adb.StreamingDenoiser(m, n_fft=n_fft, hop_length=hop, n_mels=64, sample_rate=sr, sessions=1)
    rng = np.random.default_rng(0)
    sig = (rng.standard_normal((1, n_fft + hop * (hops + 50))) * 0.1).astype(np.float32)
    lat, devt = [], []
    for i in range(hops + 50):
        t0 = time.perf_counter(); sdn.step(sig[:, i * hop: i * hop + n_fft]); dt = time.perf_counter() - t0
        if i >= 50: lat.append(dt * 1e3)
    # device time of the graph alone
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(hops):
        a.record(); sdn._graph.replay(); b.record(); torch.cuda.synchronize(); devt.append(a.elapsed_time(b))
    # host floor: an (almost) empty graph replay + sync
    g = torch.cuda.CUDAGraph(); z = torch.zeros(32, device=dev); s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        z.add_(1); s.synchronize()
        with torch.cuda.graph(g, stream=s): z.add_(1)
    torch.cuda.synchronize()
    floor = []
    for i in range(hops):
        t0 = time.perf_counter(); g.replay(); torch.cuda.current_stream().synchronize(); floor.append((time.perf_counter() - t0) * 1e3)
    # GL alone for a 3-frame hop
    from audio_denoising_b200 import _cabi
    lib = _cabi.lib(); plan = sdn.plan
    mag = torch.rand(1, 3, plan.frame_stride, device=dev); wave = torch.empty(1, plan.out_length(3), device=dev)
    ws = torch.empty(lib.b2d_griffinlim_workspace_bytes(plan.handle, 1, 3), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    glt = []
    for i in range(200):
        a.record(); _cabi.check(lib.b2d_griffinlim_frames(plan.handle, mag.data_ptr(), None, 5, 1, 3, 32, 0.99, None, wave.data_ptr(), ws.data_ptr(), ws.numel(), st)); b.record()
        torch.cuda.synchronize(); glt.append(a.elapsed_time(b))
    print(json.dumps(dict(n_fft=n_fft, step_p50_ms=round(float(np.percentile(lat, 50)), 4), graph_device_p50_ms=round(float(np.percentile(devt, 50)), 4),
                          empty_graph_replay_sync_p50_ms=round(float(np.percentile(floor, 50)), 4), gl_only_p50_ms=round(float(np.percentile(glt[20:], 50)), 4))), flush=True)
