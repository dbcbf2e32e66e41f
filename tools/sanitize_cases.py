"""A handful of SMALL cases of every tensor-core kernel family for compute-sanitizer (memcheck / racecheck /
synccheck are 10-100x slower than a plain run): im2col conv (+ residual, + fused shortcut), the CTA-pair kernel, the
window kernel (incl. temporal scatter), the fused bottleneck block, and the memory-bound ops.

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import gpu_check_fused as GF
import gpu_check_ops as G

ok = True


def report(name, info):
    global ok
    good = bool(info.get("ok"))
    ok &= good
    print(("ok   " if good else "FAIL ") + name, {k: v for k, v in info.items() if k in ("max_abs_err", "n_bad", "algo")}, flush=True)


if "--thin-only" in sys.argv:   # quick pass over the warp-MMA bottleneck kernel alone (--case NAME: one case)
    fused = {c[0]: c for c in GF.CASES}
    names = ("thin_tiny_7x7", "thin_odd_13x11_d8", "thin_rows5_grid3_d8", "thin_rows3_grid5", "thin_t1",
             "thin_pitched_out", "thin_proj_odd_13x11", "thin_proj_x16_odd")
    if "--case" in sys.argv:
        names = (sys.argv[sys.argv.index("--case") + 1],)
    for name in names:
        report("thin/" + name, GF.run_case(fused[name]))
    torch.cuda.synchronize()
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)
conv = {c[0]: c for c in G.CONV_CASES}
for name in ("pw_64_64_relu_res", "sp3_64_64", "tm3_16_16", "lat5_s4_16_16", "concat_slice"):
    report("conv/" + name, G.run_conv_case(conv[name], 0))
two = {c[0]: c for c in G.TWO_SM_CASES}
for name in ("sm2_pw_256_256_res", "sm2_sp3_512_7x7", "sm2r_pw_128_512"):
    report("two_sm/" + name, G.run_conv_case(two[name], 0))
dual = {c[0]: c for c in G.DUAL_CASES}
report("dual/dual_64_64_128_pitch", G.run_dual_case(*dual["dual_64_64_128_pitch"]))
win = {c[0]: c for c in G.WINDOW_CASES}
for name in ("w_sp3_64_64_7_wrap", "w_g_sp3_32_32_J2", "w_stem_fast_J4_tsc_T2", "w_sp33_kt3_16_res_tsc"):
    report("window/" + name, G.run_window_case(*win[name]))
fused = {c[0]: c for c in GF.CASES}
for name in ("tiny_7x7", "crop64_s4", "walk3_grid3", "pitched"):
    report("fused/" + name, GF.run_case(fused[name]))
# the warp-MMA walk kernel (algo 1): identity blocks d = 8 / 16, several strips and walks per CTA, one-frame clips, a
# projection block on 16-wide pixels
for name in ("thin_tiny_7x7", "thin_odd_13x11_d8", "thin_rows5_grid3_d8", "thin_rows3_grid5", "thin_t1", "thin_pitched_out",
             "thin_proj_odd_13x11", "thin_proj_x16_odd"):
    report("thin/" + name, GF.run_case(fused[name]))
# fused stems (kt = 1 and the temporal-scatter kt = 5 kernel) and a whole forward through a clip program
from vidsitu_b200 import ops as _ops
from vidsitu_b200.lib import VSB_BF16 as _BF16
for kt, t in ((1, 2), (5, 8)):
    n, crop = 1, 64
    g = torch.Generator().manual_seed(kt)
    fr = torch.randint(0, 256, (n, t, crop, crop, 3), dtype=torch.uint8, generator=g).cuda()
    xin = _ops.Act(torch.zeros(n * t * crop * (crop + 16) * 4, dtype=torch.bfloat16, device="cuda"), n, t, crop, crop + 16, 4, 4, c_real=3)
    _ops.pack_frames(fr, list(range(t)), [0.45] * 3, [0.225] * 3, xin, _BF16, False, 3)
    w = (torch.randn((64, 3, kt, 7, 7), generator=g) * 0.1).cuda()
    order = (0,) if kt == 1 else (0, 2, 1, 4, 3)
    q = torch.zeros((len(order), 64, 7, 8, 4), device="cuda")
    for i, k in enumerate(order):
        q[i, :, :, :7, :3] = w[:, :, k].permute(0, 2, 3, 1)
    out = _ops.Act(torch.zeros(n * t * 16 * 16 * 64, dtype=torch.bfloat16, device="cuda"), n, t, 16, 16, 64, 64)
    plan = _ops.StemPoolPlan(xin, 3, q.bfloat16().contiguous(), torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), out, crop, kt=kt)
    plan.run()
    torch.cuda.synchronize()
    x = xin.nthwc()[:, :, :, 3:3 + crop, :].permute(0, 4, 1, 2, 3).float()
    y = torch.relu(torch.nn.functional.conv3d(x, w.bfloat16().float(), None, (1, 2, 2), (kt // 2, 3, 3))).bfloat16().float()
    ref = torch.nn.functional.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1)).permute(0, 2, 3, 4, 1)
    err = float((out.buf.view(n, t, 16, 16, 64).float() - ref).abs().max())
    report(f"stem_pool/kt{kt}", {"ok": err <= 0.05 * float(ref.abs().max()), "max_abs_err": err})
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import build_model, synthetic_frames
model, cfg, _ = build_model("slow_fast_nl_r50_8x8", seed=1, crop=64)
model = model.cuda()
eng = model._engine(1, torch.device("cuda"))
eng.load_frames(synthetic_frames(1, 32, 64, seed=2).cuda())
eng.run()
torch.cuda.synchronize()
want = eng.feats.clone()
prog = eng.build_program()
eng.feats.zero_()
prog.run()
torch.cuda.synchronize()
report("program/sf50_crop64", {"ok": bool(torch.equal(eng.feats, want)), "max_abs_err": float((eng.feats - want).abs().max())})
# frame ingest: hybrid and fully-on-device JPEG decode (4:2:0, odd size; 4:4:4) against Pillow
import io as _io
import numpy as _np
from PIL import Image as _Image
from common import jpeg_bytes as _jpeg_bytes
from vidsitu_b200.jpeg import JpegBatchDecoder as _JB, JpegDecoder as _JD
_datas = [_jpeg_bytes(77, 53, 90, 2, "noisy"), _jpeg_bytes(48, 64, 75, 0, "noisy"), _jpeg_bytes(50, 70, 85, 1, "smooth")]
_out = torch.zeros((3, 32, 32, 3), dtype=torch.uint8, device="cuda")
_okb = _JB().decode_resize(_datas, list(_out))
torch.cuda.synchronize()
_ref = [_np.array(_Image.open(_io.BytesIO(d)).convert("RGB").resize((32, 32))) for d in _datas]
report("jpeg/device_batch", {"ok": all(_okb) and all(_np.array_equal(_out[i].cpu().numpy(), _ref[i]) for i in range(3))})
_o1 = torch.zeros((32, 32, 3), dtype=torch.uint8, device="cuda")
_JD(128, 128).decode_resize(_datas[0], _o1)
torch.cuda.synchronize()
report("jpeg/hybrid", {"ok": bool(_np.array_equal(_o1.cpu().numpy(), _ref[0]))})
if "--mem" in sys.argv:
    for k, v in G.run_mem_checks().items():
        report("mem/" + k, v)
torch.cuda.synchronize()
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
