"""Role timeline of the tensor-core conv kernels on SF50 layers (VSB_WIN_DEBUG counters).
    python tools/win_debug.py [op-substring ...]"""
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
names = ["prod_wait_empty", "mma_wait_acc", "mma_wait_data", "mma_issue", "epi_wait_acc", "epi_wait_slab",
         "epi_math", "epi_wait_store_read", "cta_total", "tiles", "epi_fence"]
lib = L.load()
for name, fn, _ in eng.trunk_ops:
    if not any(w in name for w in want):
        continue
    plan = getattr(fn, "__self__", None)
    if plan is None or not hasattr(plan, "_h"):
        continue
    out = (C.c_longlong * 16)()
    if lib.vsb_debug_conv_stats(plan._h, out) != 0:
        print(name, "-> no counters"); continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.check(lib.vsb_debug_conv_stats(plan._h, out), "stats")
    v = list(out)
    grid, stages, smem = v[13], v[14], v[15]
    tiles = max(v[9], 1)          # tiles seen by epilogue warp 0 of every CTA = all tiles
    per_cta = tiles / grid
    print(f"{name}: {e0.elapsed_time(e1)*1e3:.1f} us  grid={grid} stages={stages} smem={smem//1024}K tiles={tiles} "
          f"({per_cta:.1f}/CTA)  cta_total={v[8]/grid:.0f} clk = {v[8]/grid/per_cta:.0f} clk/tile")
    print("   " + "  ".join(f"{names[k]}={v[k]/tiles:.0f}" for k in (0, 1, 2, 3, 4, 5, 6, 7)))
