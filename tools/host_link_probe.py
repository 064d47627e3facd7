"""Aggregate host <-> device copy bandwidth with N ranks at once (torchrun), default pinned memory vs write-combined pinned
memory for the upload source, each direction alone and both together.  Explains the e2e ceiling at N = 8."""
import ctypes, json, os, sys, time
import torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cudart = ctypes.CDLL("libcudart.so")
NBYTES = 256 * 64000 * 2  # one int16 batch


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.int16)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(src_host, dst_host, up, down, reps=40):
    d_in = torch.empty(NBYTES // 2, dtype=torch.int16, device=dev)
    d_out = torch.empty(NBYTES // 2, dtype=torch.int16, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(src_host, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                dst_host.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = reps * NBYTES * (int(up) + int(down)) / dt / 1e9
    t = torch.tensor([gbs], device=dev)
    if world > 1:
        dist.all_reduce(t)
    return round(float(t.item()), 1)


try:
    wc = host_alloc(NBYTES, 4)  # cudaHostAllocWriteCombined
    wc.fill_(1)
except Exception as e:  # noqa
    wc = None
plain_in = torch.ones(NBYTES // 2, dtype=torch.int16).pin_memory()
plain_out = torch.empty(NBYTES // 2, dtype=torch.int16).pin_memory()
# zero-copy download: the float -> int16 kernel of the library writes straight into pinned host memory (SM stores over PCIe
# instead of the copy engine)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_denoising_b200 import _cabi


def run_zero_copy(dst_host, src_host=None, reps=40):
    lib = _cabi.lib()
    f = torch.rand(NBYTES // 2, device=dev) - 0.5
    d_in = torch.empty(NBYTES // 2, dtype=torch.int16, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if src_host is not None:
            with torch.cuda.stream(s1):
                d_in.copy_(src_host, non_blocking=True)
        with torch.cuda.stream(s2):
            _cabi.check(lib.b2d_float_to_pcm16(f.data_ptr(), f.numel(), dst_host.data_ptr(), s2.cuda_stream))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = reps * NBYTES * (2 if src_host is not None else 1) / dt / 1e9
    t = torch.tensor([gbs], device=dev)
    if world > 1:
        dist.all_reduce(t)
    return round(float(t.item()), 1)


rows = {}
for name, src in (("pinned", plain_in), ("write_combined", wc)):
    if src is None:
        continue
    assert src.is_pinned()
    rows[name] = dict(h2d_only=run(src, plain_out, True, False), d2h_only=run(src, plain_out, False, True), both=run(src, plain_out, True, True))
rows["zero_copy_download_kernel"] = dict(d2h_only=run_zero_copy(plain_out), both=run_zero_copy(plain_out, plain_in))
if rank == 0:
    print(json.dumps(dict(n_gpus=world, aggregate_GBps=rows)))
if world > 1:
    dist.destroy_process_group()
