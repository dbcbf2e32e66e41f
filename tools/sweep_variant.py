"""Per-launch timings of the SF50 batch-N engine under one tuning variant (run once per variant: some
knobs are environment variables read once per process).

    VSB_EPI_WARPS=16 python tools/sweep_variant.py NAME '{"*": {"block_n": 128}}'  -> gpurun_out/sweep_NAME.json
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import build_model, synthetic_frames

name = sys.argv[1]
tune = json.loads(sys.argv[2]) if len(sys.argv) > 2 else None
n = int(os.environ.get("VSB_PROFILE_CLIPS", "64"))
mdl = os.environ.get("VSB_PROFILE_MODEL", "slow_fast_nl_r50_8x8")
model, cfg, _ = build_model(mdl, seed=0, crop=224, micro_batch=n)
model.tune = tune
model = model.cuda()
eng = model._engine(n, torch.device("cuda"))
frames = synthetic_frames(n, cfg.sf_mdl.DATA.NUM_FRAMES, 224, seed=1).cuda()
eng.load_frames(frames)
eng.run()
torch.cuda.synchronize()
ops = eng.time_ops(iters=5)
eng.capture()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    eng.replay()
e0.record()
for _ in range(10):
    eng.replay()
e1.record()
torch.cuda.synchronize()
step = e0.elapsed_time(e1) / 10
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"name": name, "tune": tune, "step_ms": step, "ops": [[o, round(ms, 4)] for o, ms, _ in ops]},
          open(f"gpurun_out/sweep_{name}.json", "w"))
chk = float(eng.feats.double().abs().sum())   # bit-exact across scheduling variants
print(name, "step_ms", round(step, 3), "sum_ops", round(sum(ms for _, ms, _ in ops), 3), "feats_abs_sum", repr(chk), flush=True)
