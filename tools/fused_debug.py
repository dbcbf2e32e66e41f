"""Role timeline of the fused bottleneck kernel on the full-size Fast-pathway blocks (VSB_FUSED_DEBUG=1):
average cycles per CTA each role spends waiting / working, and cycles per tile.
    VSB_FUSED_DEBUG=1 python tools/fused_debug.py [case ...] [--tune k=v ...]"""
import os
import sys

os.environ.setdefault("VSB_FUSED_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch

from gpu_check_fused import BENCH_CASES, make_case
from vidsitu_b200.ops import Act, BottleneckPlan

NAMES = ["prod:wait_slot", "mma:wait_aacc", "mma:wait_x", "mma:wait_aring", "mma:wait_bacc", "mma:wait_P", "mma:wait_cacc",
         "aepi:wait", "aepi:work", "bepi:wait", "bepi:work", "cepi:wait_full", "cepi:wait_slab", "cepi:work",
         "cepi:wait_read", "cta_total", "a_tiles", "c_tiles_w8", "mma:issue_a", "mma:issue_b", "mma:issue_c", "mmaA:wait_ringfree"]


def main():
    args = [a for a in sys.argv[1:] if "=" not in a]
    tune = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:] if "=" in a)}
    for case in BENCH_CASES:
        if args and case[0] not in args:
            continue
        name, n, t, h, w, c, d, kt, _, _, _ = case
        xbuf, wa, wb, wc, aff, outbuf, xp, op = make_case(case)
        plan = BottleneckPlan(Act(xbuf, n, t, h, w, c, xp), Act(outbuf, n, t, h, w, c, op), d, kt,
                              wa.permute(0, 2, 3, 4, 1).reshape(d, kt, c).contiguous(),
                              wb.permute(0, 2, 3, 4, 1).reshape(d, 9, d).contiguous(), wc.reshape(c, d).contiguous(), *aff,
                              **tune)
        info = plan.info()
        for _ in range(2):
            plan.run()
        plan.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run()
        e1.record()
        e1.synchronize()
        st = plan.stats()
        grid = info["grid"]
        tiles = st[16] / grid
        print(f"== {name} {tune} plan={info} ms={e0.elapsed_time(e1):.3f} a_tiles/CTA={tiles:.1f} "
              f"clk/tile={st[15] / grid / max(tiles, 1):.0f}")
        for i, nm in enumerate(NAMES):
            if i in (16, 17):
                continue
            print(f"   {nm:16s} {st[i] / grid:12.0f} clk/CTA  {st[i] / max(st[16], 1):8.0f} /tile")


if __name__ == "__main__":
    main()
