"""Role timeline of the window conv kernel on the SF50 layers that use it (VSB_WIN_DEBUG counters).
    VSB_WIN_DEBUG=1 python tools/win_debug.py [op-substring ...]"""
import ctypes as C, os, sys
os.environ["VSB_WIN_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import build_model, synthetic_frames
from vidsitu_b200 import lib as L

n = int(os.environ.get("VSB_PROFILE_CLIPS", "64"))
model, cfg, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=224, micro_batch=n)
model = model.cuda()
eng = model._engine(n, torch.device("cuda"))
frames = synthetic_frames(n, 32, 224, seed=1).cuda()
eng.load_frames(frames); eng.run(); torch.cuda.synchronize()
want = sys.argv[1:] or ["s2.pathway0_res1.branch2.b", "s2.pathway1_res1.branch2.b", "s3.pathway1_res1.branch2.b"]
names = ["prod_wait_empty", "mma_wait_acc", "mma_wait_window", "mma_issue", "epi_wait_acc", "epi_wait_slab",
         "epi_math", "epi_wait_store_read", "cta_total", "tiles", "epi_fence_syncwarp", "epi_tmem_ld", "x"]
lib = L.load()
plans = {}
for pl in eng._keep:
    if hasattr(pl, "_h"):
        plans[id(pl)] = pl
for name, fn, _ in eng.trunk_ops:
    if not any(w in name for w in want):
        continue
    plan = getattr(fn, "__self__", None)
    if plan is None:
        continue
    out = (C.c_longlong * 16)()
    if lib.vsb_debug_conv_stats(plan._h, out) != 0:
        print(name, "-> not a window plan"); continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.check(lib.vsb_debug_conv_stats(plan._h, out), "stats")
    v = list(out)[:13]
    tiles = max(v[9], 1)
    ctas = 148
    print(f"{name}: {e0.elapsed_time(e1)*1e3:.1f} us, tiles(epi warp0)={tiles}, cta_total/CTA={v[8]/ctas:.0f} clk")
    for k in (0, 1, 2, 3, 4, 5, 6, 10, 7):
        print(f"   {names[k]:22s} {v[k]/tiles:9.0f} clk/tile")
