"""TMA load-path throughput on B200 (vsb_debug_tma_rate): bytes per cycle per SM for the box shapes the kernels use.
-> gpurun_out/tma_rate.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vidsitu_b200 import lib as L

lib = L.load()
grid = 148
clk = torch.zeros(grid, dtype=torch.int64, device="cuda")
out = []
# (label, n, t, h, w, c)
shapes = [("fast_res3 28x28x64", 64, 32, 28, 28, 64), ("fast_res4 14x14x128", 64, 32, 14, 14, 128),
          ("fast_res2 56x28x64 (2-px groups)", 64, 32, 56, 28, 64), ("slow_res2 56x56x256", 16, 8, 56, 56, 256)]
for label, n, t, h, w, c in shapes:
    x = torch.randn((n, t, h, w, c), device="cuda").bfloat16()
    kc = 64
    chunk = 128 * kc * 2
    for mode, kt, box_rows, stages in ((1, 1, 1, 6), (1, 1, 1, 12), (1, 3, 1, 6), (0, 1, 1, 6), (0, 1, 1, 12), (0, 3, 1, 6), (0, 3, 1, 12),
                                       (0, 3, 2, 6), (0, 3, 4, 6), (0, 3, 8, 6), (2, 1, 1, 6), (2, 3, 1, 6), (2, 3, 1, 12)):
        rp = 8
        while rp < w + 1:
            rp *= 2
        if (128 // rp) % box_rows:
            continue
        tiles = 200
        for _ in range(2):
            L.check(lib.vsb_debug_tma_rate(x.data_ptr(), n, t, h, w, c, mode, kt, box_rows, stages, tiles, grid, clk.data_ptr(), None), "tma_rate")
            torch.cuda.synchronize()
        c_avg = float(clk.float().mean())
        nbytes = tiles * kt * (c // kc) * chunk
        rec = dict(shape=label, mode={0: "5d tiled [kc,RP,rows]", 1: "2d [kc,128]", 2: "5d im2col, 128 padded positions"}[mode], kt=kt, box_rows=box_rows, stages=stages,
                   bytes_per_clk_per_sm=round(nbytes / c_avg, 1), clk_per_16KB=round(c_avg / (nbytes / chunk), 0),
                   chip_TBs_at_1p85GHz=round(nbytes / c_avg * grid * 1.85e9 / 1e12, 2))
        out.append(rec)
        print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tma_rate.json", "w"), indent=1)
