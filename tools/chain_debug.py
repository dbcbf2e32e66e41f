"""Timing of the chained launch pairs (tile_signal / tile_wait): producer alone, consumer alone, pair."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import build_model, synthetic_frames

os.environ.setdefault("VSB_CHAIN", "1")
clips = int(os.environ.get("CLIPS", "64"))
model, cfg, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=224, micro_batch=clips)
model = model.cuda()
eng = model._engine(clips, torch.device("cuda"))
frames = synthetic_frames(clips, cfg.sf_mdl.DATA.NUM_FRAMES, 224, seed=1).cuda()
eng.load_frames(frames)
eng.run()
torch.cuda.synchronize()


def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for c_name, a_name, c_run, a_run, c_plan in eng.chained:
    info = c_plan.info()
    pair = t(lambda: (c_run(), a_run()))
    # producer alone (counters keep growing: harmless), then consumer alone on raised counters
    c_alone = t(c_run)
    def a_alone_fn():
        pass
    best_a = 1e9
    for _ in range(5):
        c_run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a_run(); e1.record(); e1.synchronize()
        best_a = min(best_a, e0.elapsed_time(e1))
    print(f"{c_name} + {a_name}: pair {pair:.3f} ms, producer alone {c_alone:.3f}, consumer alone {best_a:.3f}; producer plan {info}")
