"""Runs selected launches of the SF50 batch-N engine between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full --import-source on -o gpurun_out/prof python tools/profile_ops.py OP [OP...]
OP is a substring of the op name (e.g. s2.pathway0_res1.branch2.b); 'all' profiles one whole step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import build_model, synthetic_frames

n = int(os.environ.get("VSB_PROFILE_CLIPS", "64"))
model, cfg, _ = build_model(os.environ.get("VSB_PROFILE_MODEL", "slow_fast_nl_r50_8x8"), seed=0, crop=224, micro_batch=n)
model = model.cuda()
eng = model._engine(n, torch.device("cuda"))
frames = synthetic_frames(n, cfg.sf_mdl.DATA.NUM_FRAMES, 224, seed=1).cuda()
eng.load_frames(frames)
eng.run()
torch.cuda.synchronize()
want = sys.argv[1:] or ["all"]
ops = eng.trunk_ops + eng.head_ops
torch.cuda.profiler.start()
if want == ["all"]:
    eng.load_frames(frames)
    eng.run()
else:
    for name, fn, _ in ops:
        if any(w == name or (w.endswith("*") and name.startswith(w[:-1])) for w in want):
            print("profiling", name, flush=True)
            fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
