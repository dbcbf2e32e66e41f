"""BASELINE.json configs[3] (C4): SlowFast-R101 16x8 feature extraction, videos sharded over the ranks of one
box, NCCL all-gather of the [n, 2304] features; the gathered rows must equal a single-GPU run of all the
videos BIT FOR BIT (same kernels, per-clip math independent of the batch).  Also times the sharded run.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gpu_shard_check.py [--videos 8] [--model ...]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from common import build_model, synthetic_frames
from vidsitu_b200.dist import gather_rows, shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="slow_fast_r101_16x8")
ap.add_argument("--videos", type=int, default=8)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
model, cfg, _ = build_model(args.model, seed=0, crop=224, micro_batch=20)
model = model.to(dev)
t = cfg.sf_mdl.DATA.NUM_FRAMES
frames = synthetic_frames(5 * args.videos, t, 224, seed=4321)          # the same on every rank (seeded CPU generator)
lo, hi = shard_range(args.videos, rank, world)
mine = frames[5 * lo:5 * hi].to(dev)
feats_local = model.extract_features(mine)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
e0.record()
feats_local = model.extract_features(mine)
allf = gather_rows(feats_local, 5 * args.videos)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    single = model.extract_features(frames.to(dev))
    same = bool(torch.equal(single, allf))
    print(json.dumps({"check": "sharded == single-GPU, bit for bit", "model": args.model, "world": world,
                      "clips": 5 * args.videos, "equal": same, "max_abs_diff": float((single - allf).abs().max()),
                      "sharded_ms": round(float(ms.item()), 3), "clips_per_s": round(5 * args.videos / float(ms.item()) * 1e3, 1)}))
    assert same
if world > 1:
    dist.destroy_process_group()
