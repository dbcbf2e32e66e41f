"""Per-launch times of the ops whose name contains FILTER, under tune {"*": {KNOB: value}} for each value."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import build_model, synthetic_frames

knob, values, flt = sys.argv[1], sys.argv[2].split(","), sys.argv[3]
clips = int(os.environ.get("CLIPS", "64"))
name = os.environ.get("MODEL", "slow_fast_nl_r50_8x8")
res = {}
for v in values:
    val = {"0": False, "1": True}.get(v, v)
    model, cfg, _ = build_model(name, seed=0, crop=224, micro_batch=clips, tune={"*": {knob: val}})
    model = model.cuda()
    eng = model._engine(clips, torch.device("cuda"))
    frames = synthetic_frames(clips, cfg.sf_mdl.DATA.NUM_FRAMES, 224, seed=1).cuda()
    eng.load_frames(frames)
    eng.run()
    torch.cuda.synchronize()
    res[v] = {nm: ms for nm, ms, _ in eng.time_ops(5) if flt in nm}
    del eng, model
    torch.cuda.empty_cache()
names = list(res[values[0]])
tot = {v: 0.0 for v in values}
for nm in names:
    row = [res[v].get(nm, float("nan")) for v in values]
    for v, r in zip(values, row):
        tot[v] += r if r == r else 0
    print(f"{nm:40s} " + "  ".join(f"{r:.3f}" for r in row))
print("total", tot)
