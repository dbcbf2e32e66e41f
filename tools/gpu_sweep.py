"""BASELINE.json configs[4] (C5): SlowFast-R50 8x8 batch sweep, N = 1 .. 512 event clips, on 1 / 2 / 4 / 8 GPUs.

    python tools/gpu_sweep.py                                   # 1 GPU
    torchrun --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 tools/gpu_sweep.py

N clips are sharded over the ranks in contiguous blocks (vidsitu_b200.dist.shard_range on clips), every rank runs its
shard in micro-batches of <= 64 clips (uint8 frames already in HBM: pack + graph-replayed forward), the features are
all-gathered, and the line reports N / (max over ranks of the device time).  Appends one JSON record per N to
gpurun_out/sweep_w{W}.json.  The host-CPU baseline of the same forward is bench.py's `cpu_baseline`.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from common import build_model, synthetic_frames
from vidsitu_b200.dist import shard_range

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
MB = 64
model, cfg, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=224, micro_batch=MB)
model = model.to(dev)
pool = synthetic_frames(MB, 32, 224, seed=99 + rank).to(dev)        # distinct clips per rank; re-used across micro-batches
records = []
for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    lo, hi = shard_range(n, rank, world) if n >= world else ((rank, rank + 1) if rank < n else (0, 0))
    mine = hi - lo
    chunks = [min(MB, mine - s) for s in range(0, mine, MB)]
    feats_all = torch.zeros((max(mine, 1), 2304), device=dev)

    def run():
        off = 0
        for c in chunks:
            f = model.extract_features(pool[:c])
            feats_all[off:off + c] = f
            off += c
        if world > 1:
            # equal-size gather of the longest shard (the ragged tail is padding), as bench.py does each step
            longest = max(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)) if n >= world else 1
            buf = torch.zeros((world * longest, 2304), device=dev)
            dist.all_gather_into_tensor(buf, torch.nn.functional.pad(feats_all[:mine], (0, 0, 0, longest - mine))
                                        if mine < longest else feats_all[:longest].contiguous())

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = 10 if n <= 64 else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = {"clips": n, "gpus": world, "ms": round(float(ms.item()), 3), "clips_per_s": round(n / float(ms.item()) * 1e3, 1),
           "clips_on_rank0": mine if rank == 0 else None, "micro_batches_rank0": chunks if rank == 0 else None}
    if rank == 0:
        records.append(rec)
        print(json.dumps(rec), flush=True)
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"model": "slow_fast_nl_r50_8x8", "dtype": "bf16", "gpus": world, "timing": "CUDA events, max over ranks, "
               "device-resident uint8 frames, pack + CUDA-graph forward + feature all-gather", "records": records},
              open(os.path.join(ROOT, "gpurun_out", f"sweep_w{world}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
