"""One launch of a full-size fused bottleneck block for `ncu --set full` (profile_range)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gpu_check_fused import BENCH_CASES, make_case
from vidsitu_b200.ops import Act, BottleneckPlan
want = sys.argv[1] if len(sys.argv) > 1 else "bench_fast_s3"
for case in BENCH_CASES:
    if case[0] != want:
        continue
    name, n, t, h, w, c, d, kt, _, _, _ = case
    xbuf, wa, wb, wc, aff, outbuf, xp, op = make_case(case)
    plan = BottleneckPlan(Act(xbuf, n, t, h, w, c, xp), Act(outbuf, n, t, h, w, c, op), d, kt,
                          wa.permute(0, 2, 3, 4, 1).reshape(d, kt, c).contiguous(),
                          wb.permute(0, 2, 3, 4, 1).reshape(d, 9, d).contiguous(), wc.reshape(c, d).contiguous(), *aff)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
