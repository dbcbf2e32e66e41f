"""tcgen05.mma / tcgen05.commit pacing on B200: cycles per MMA as a function of the commit frequency, the operand row
pitch and a row-shifted start address (vsb_debug_umma_rate2).  -> gpurun_out/umma_rate2.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vidsitu_b200 import lib as L

lib = L.load()
grid = 148
clk = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
out = []
for n in (16, 32, 64):
    for row_bytes in (128, 64, 32):
        for shift, walk in ((0, 0), (1, 0), (1, 1), (8, 0)):
            for per_group, commit in ((9, 0), (9, 1), (1, 1), (3, 1), (24, 1)):
                groups = 1800 // per_group
                L.check(lib.vsb_debug_umma_rate2(n, groups, per_group, commit, row_bytes, shift, walk, grid,
                                                 clk.data_ptr(), None), "rate2")
                torch.cuda.synchronize()
                c = clk.view(grid, 2).float().mean(0)
                tot = groups * per_group
                rec = dict(n=n, row_bytes=row_bytes, shift=shift, walk=walk, per_group=per_group, commit=commit,
                           issue_clk_per_mma=round(float(c[0]) / tot, 1), done_clk_per_mma=round(float(c[1]) / tot, 1),
                           done_clk_per_group=round(float(c[1]) / groups, 1))
                out.append(rec)
                print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/umma_rate2.json", "w"), indent=1)
