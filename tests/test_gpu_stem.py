"""Fused [1,7,7] stem (vsb_stem_pool_*, csrc/stem_pool_sm100.cu): conv + frozen BN + ReLU + 1x3x3/s2 max-pool in
one kernel, against plain torch fp32 on the same bf16-rounded operands (stem_helper.py:157-178) and against the
unfused two-launch path of the engine."""
import numpy as np
import pytest
import torch

from common import build_model, synthetic_frames

pytestmark = pytest.mark.gpu


def _torch_stem(x_bf16_ncthw, w, scale, bias):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = torch.nn.functional.conv3d(x_bf16_ncthw.float(), w.bfloat16().float(), None, (1, 2, 2), (w.shape[2] // 2, 3, 3))
    y = torch.relu(y * scale.view(1, -1, 1, 1, 1) + bias.view(1, -1, 1, 1, 1)).bfloat16().float()
    return torch.nn.functional.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1))


@pytest.mark.parametrize("crop,n,t,pitch,kt", [(64, 2, 3, 64, 1), (224, 1, 2, 80, 1), (96, 3, 1, 72, 1),
                                               (64, 2, 8, 64, 5), (96, 3, 3, 64, 5), (64, 1, 16, 72, 5), (224, 2, 8, 64, 5),
                                               (224, 8, 8, 80, 5), (224, 24, 1, 64, 1)])   # several runs / tiles per CTA
def test_fused_stem_matches_torch(crop, n, t, pitch, kt):
    from vidsitu_b200 import ops
    from vidsitu_b200.lib import VSB_BF16
    from vidsitu_b200.ops import Act

    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(crop + n)
    frames = torch.randint(0, 256, (n, t, crop, crop, 3), dtype=torch.uint8, generator=g).to(dev)
    w = (torch.randn((64, 3, kt, 7, 7), generator=g) * 0.1).to(dev)
    scale = (torch.rand(64, generator=g) + 0.5).to(dev)
    bias = (torch.randn(64, generator=g) * 0.3).to(dev)
    mean, std = [0.45, 0.45, 0.45], [0.225, 0.225, 0.225]
    w_buf = crop + 16
    xin = Act(torch.zeros(n * t * crop * w_buf * 4, dtype=torch.bfloat16, device=dev), n, t, crop, w_buf, 4, 4, c_real=3)
    ops.pack_frames(frames, list(range(t)), mean, std, xin, VSB_BF16, False, 3)
    order = (0,) if kt == 1 else (0, 2, 1, 4, 3)
    q = torch.zeros((len(order), 64, 7, 8, 4), dtype=torch.float32, device=dev)
    for i, k in enumerate(order):
        q[i, :, :, :7, :3] = w[:, :, k].permute(0, 2, 3, 1)
    wq = q.bfloat16().contiguous()
    po = crop // 4
    # the untouched channel slice [64, pitch) must survive (it belongs to the lateral connection)
    out_buf = torch.full((n * t * po * po * pitch,), 7.0, dtype=torch.bfloat16, device=dev)
    out = Act(out_buf, n, t, po, po, 64, pitch)
    plan = ops.StemPoolPlan(xin, 3, wq, scale, bias, out, crop, kt=kt)
    plan.run()
    plan.run()     # idempotent: the output is zero-filled inside the run
    torch.cuda.synchronize()
    got = out_buf.view(n, t, po, po, pitch).float()
    x = xin.nthwc()[:, :, :, 3:3 + crop, :].permute(0, 4, 1, 2, 3).contiguous()     # [n, 3, t, h, w] bf16-rounded input
    ref = _torch_stem(x, w, scale, bias).permute(0, 2, 3, 4, 1)                      # [n, t, po, po, 64]
    if pitch > 64:
        assert torch.all(got[..., 64:] == 7.0)
    err = (got[..., :64] - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + 1e-3          # one bf16 ulp (fp32 accumulation order differs)
    assert bool((err <= tol).all()), float((err - tol).max())
    assert float((err > 0).float().mean()) < 0.02


@pytest.mark.parametrize("name", ["slow_fast_nl_r50_8x8", "i3d_r50_8x8"])
def test_engine_with_fused_stem_matches_unfused_engine(name):
    feats = {}
    for fuse in (True, False):
        model, cfg, _ = build_model(name, seed=5, crop=64, tune={"*": {"fuse_stem": fuse}})
        model = model.cuda()
        eng = model._engine(2, torch.device("cuda"))
        assert (len(eng.fused_stems) == 1) == fuse
        frames = synthetic_frames(2, cfg.sf_mdl.DATA.NUM_FRAMES, 64, seed=77).cuda()
        eng.load_frames(frames)
        eng.replay()
        torch.cuda.synchronize()
        feats[fuse] = eng.feats.cpu().numpy().copy()
    a, b = feats[True], feats[False]
    cos = float(((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))).min())
    assert cos > 0.99999 and float(np.abs(a - b).max() / np.abs(b).max()) < 3e-3, cos
