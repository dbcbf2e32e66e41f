"""One full-size launch of the fused stem kernel (kt = 1: SlowFast slow stem, 64 clips x 8 frames; kt = 5: I3D stem)
inside a profiler range, for `ncu --set full --profile-from-start off`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vidsitu_b200 import ops
from vidsitu_b200.lib import VSB_BF16
from vidsitu_b200.ops import Act

dev = torch.device("cuda")
n, t, crop = 64, 8, 224
w_buf = crop + 16
g = torch.Generator().manual_seed(0)
frames = torch.randint(0, 256, (n, t, crop, crop, 3), dtype=torch.uint8, generator=g).to(dev)
xin = Act(torch.zeros(n * t * crop * w_buf * 4, dtype=torch.bfloat16, device=dev), n, t, crop, w_buf, 4, 4, c_real=3)
ops.pack_frames(frames, list(range(t)), [0.45] * 3, [0.225] * 3, xin, VSB_BF16, False, 3)
for kt in (1, 5):
    wq = (torch.randn((kt * 64, 7, 8, 4), generator=g) * 0.1).to(dev).bfloat16().contiguous()
    scale, bias = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    out = Act(torch.zeros(n * t * 56 * 56 * 80, dtype=torch.bfloat16, device=dev), n, t, 56, 56, 64, 80)
    plan = ops.StemPoolPlan(xin, 3, wq, scale, bias, out, crop, kt=kt)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
