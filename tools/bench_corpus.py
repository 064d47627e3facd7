"""BASELINE config 4: offline corpus denoise -- N synthetic 10 s clips sharded data-parallel over the GPUs of one box.

Launch: python tools/bench_corpus.py [--clips 10000]            (1 GPU)
        torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_corpus.py --clips 10000
Each rank takes a contiguous block of clips (sharding.shard_range), streams it through DenoisePipeline.denoise_host in
batches from pinned host memory (uploads / downloads overlapped with compute), and the job time is the max over ranks.
Host synthesis of the corpus is excluded from the timing (one 256-clip block is synthesised and re-used per rank).
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import audio_denoising_b200 as adb
from audio_denoising_b200.sharding import shard_range, max_over_ranks, gather_counts
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--clips", type=int, default=10000); ap.add_argument("--batch", type=int, default=128)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
from audio_denoising_b200.sharding import bind_host_to_gpu
numa = bind_host_to_gpu(local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
sd, cfg = bench.load_model_weights()
m = adb.GRUUNet2(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
pipe = adb.DenoisePipeline(m, n_fft=1024, hop_length=512, n_mels=64, sample_rate=16000, n_iter=32)
L = 160000
lo, hi = shard_range(args.clips, world, rank)
Bt = args.batch
block = bench.synth_batch(Bt, L, seed=77 + rank).pin_memory()
outs = [torch.empty((Bt, pipe.out_length(L)), dtype=torch.float32, pin_memory=True) for _ in range(2)]
for i in range(2): pipe.denoise_host(block, outs[i % 2], wait=False)
pipe.host_synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
done, i = 0, 0
while done < hi - lo:
    nb = min(Bt, hi - lo - done)
    pipe.denoise_host(block[:nb], outs[i % 2][:nb], wait=False)
    done += nb; i += 1
pipe.host_synchronize()
sec = time.perf_counter() - t0
worst = max_over_ranks(sec); total = gather_counts(done)
if rank == 0:
    print(json.dumps(dict(config="corpus 10 s clips, 16 kHz, n_fft 1024, GL 32 it", clips=total, n_gpus=world, seconds=round(worst, 3),
                          audio_s_per_s=round(total * 10 / worst, 1), batch=Bt, includes="pinned host->device->host for every clip")), flush=True)
if world > 1: dist.destroy_process_group()
