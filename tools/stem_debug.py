"""Timing of the stem ops of the headline config: fused conv+pool kernel vs separate conv and pool launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import build_model, synthetic_frames

clips = int(os.environ.get("CLIPS", "64"))
name = os.environ.get("MODEL", "slow_fast_nl_r50_8x8")
for fuse in (False, True):
    model, cfg, _ = build_model(name, seed=0, crop=224, micro_batch=clips, tune={"*": {"fuse_stem": fuse}})
    model = model.cuda()
    eng = model._engine(clips, torch.device("cuda"))
    frames = synthetic_frames(clips, cfg.sf_mdl.DATA.NUM_FRAMES, 224, seed=1).cuda()
    eng.load_frames(frames)
    eng.run()
    torch.cuda.synchronize()
    for nm, ms, fl in eng.time_ops(5):
        if "stem" in nm:
            print(f"fuse_stem={fuse}  {nm:40s} {ms:.3f} ms  {fl / ms / 1e9 if ms else 0:.0f} TFLOP/s")
    del eng, model
    torch.cuda.empty_cache()
