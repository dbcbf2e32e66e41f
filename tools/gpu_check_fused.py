"""GPU checks of the fused identity-bottleneck kernel (vsb_bottleneck_*, bottleneck_fused_sm100.cu), gpurun style:
every case runs in a restartable child (a trapped kernel poisons the context), results go to a JSON report.

    python tools/gpu_check_fused.py [--out gpurun_out/check_fused.json] [--bench]

The checker is plain torch on the same GPU (fp32, TF32 off) with the block's two intermediate tensors rounded to
bf16 exactly where the three-launch path rounds them -- test infrastructure only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import traceback
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from vidsitu_b200.ops import Act, BottleneckPlan

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# (name, n, t, h, w, c, d, kt, tuning, x_pitch, out_pitch)
CASES = [
    ("s3fast_like", 2, 8, 28, 28, 64, 16, 3, {}, 0, 0),
    ("s4fast_like", 2, 8, 14, 14, 128, 32, 3, {}, 0, 0),
    ("s2fast_grouped_like", 2, 8, 56, 28, 64, 16, 3, {}, 0, 0),
    ("s2fast_plain_like", 1, 8, 56, 56, 32, 16, 3, {}, 0, 0),
    ("kt1_c128_d32", 2, 4, 28, 28, 128, 32, 1, {}, 0, 0),
    ("kt1_c64_d64", 2, 4, 14, 14, 64, 64, 1, {}, 0, 0),
    ("tiny_7x7", 3, 8, 7, 7, 64, 16, 3, {}, 0, 0),
    ("crop64_s2", 2, 32, 16, 16, 32, 16, 3, {}, 0, 0),
    ("crop64_s3", 2, 32, 8, 8, 64, 16, 3, {}, 0, 0),
    ("crop64_s4", 2, 32, 4, 4, 128, 32, 3, {}, 0, 0),
    ("odd_13x11", 2, 5, 13, 11, 64, 32, 3, {}, 0, 0),
    ("walk2", 2, 8, 28, 28, 64, 16, 3, dict(walk_len=2), 0, 0),
    ("walk3_grid3", 2, 8, 28, 28, 64, 16, 3, dict(walk_len=3, grid=3), 0, 0),
    ("walk1_grid2", 1, 4, 14, 14, 128, 32, 3, dict(walk_len=1, grid=2), 0, 0),
    ("stages2", 2, 8, 14, 14, 128, 32, 3, dict(stages=2), 0, 0),
    ("pitched", 2, 4, 14, 14, 64, 16, 3, {}, 96, 80),
    ("c256_d64_kt1_small", 1, 2, 14, 14, 256, 64, 1, {}, 0, 0),
    ("many_tiles", 8, 32, 28, 28, 64, 16, 3, {}, 0, 0),
    # algo 1: the warp-MMA walk kernel (bottleneck_thin_sm100.cu), ungrouped thin blocks d = 8 / 16, c = 4 d
    ("thin_s2_like", 2, 8, 56, 56, 32, 8, 3, dict(algo=1), 0, 0),
    ("thin_s3_like", 2, 8, 28, 28, 64, 16, 3, dict(algo=1), 0, 0),
    ("thin_tiny_7x7", 3, 8, 7, 7, 64, 16, 3, dict(algo=1), 0, 0),
    ("thin_odd_13x11_d8", 2, 5, 13, 11, 32, 8, 3, dict(algo=1), 0, 0),
    ("thin_odd_13x11_d16", 2, 5, 13, 11, 64, 16, 3, dict(algo=1), 0, 0),
    ("thin_kt1_d8", 2, 4, 28, 28, 32, 8, 1, dict(algo=1), 0, 0),
    ("thin_kt1_d16", 2, 4, 14, 14, 64, 16, 1, dict(algo=1), 0, 0),
    ("thin_t1", 2, 1, 14, 14, 32, 8, 3, dict(algo=1), 0, 0),
    ("thin_t2", 2, 2, 14, 14, 64, 16, 3, dict(algo=1), 0, 0),
    ("thin_rows3_grid5", 2, 8, 28, 28, 64, 16, 3, dict(algo=1, walk_len=3, grid=5), 0, 0),
    ("thin_rows5_grid3_d8", 3, 6, 28, 28, 32, 8, 3, dict(algo=1, walk_len=5, grid=3), 0, 0),
    ("thin_grid7_d8", 3, 6, 28, 28, 32, 8, 3, dict(algo=1, walk_len=7, grid=7), 0, 0),
    ("thin_slots4", 2, 8, 28, 28, 64, 16, 3, dict(algo=1, stages=4), 0, 0),
    ("thin_pitched_out", 2, 4, 14, 14, 64, 16, 3, dict(algo=1), 0, 80),
    ("thin_crop64_s2", 2, 32, 16, 16, 32, 8, 3, dict(algo=1), 0, 0),
    ("thin_crop64_s3", 2, 32, 8, 8, 64, 16, 3, dict(algo=1), 0, 0),
    ("thin_many_steps", 8, 32, 28, 28, 64, 16, 3, dict(algo=1), 0, 0),
    # algo 1, block 0 of a stage with a projection shortcut (cin = d = 8, out 32 channels)
    ("thin_proj_s2_like", 2, 8, 56, 56, 32, 8, 3, dict(algo=1, cin=8), 0, 0),
    ("thin_proj_odd_13x11", 2, 5, 13, 11, 32, 8, 3, dict(algo=1, cin=8), 0, 0),
    ("thin_proj_kt1", 2, 4, 28, 28, 32, 8, 1, dict(algo=1, cin=8), 0, 0),
    ("thin_proj_rows5_grid3", 3, 6, 28, 28, 32, 8, 3, dict(algo=1, cin=8, walk_len=5, grid=3), 0, 0),
    ("thin_proj_pitched_out", 2, 4, 16, 16, 32, 8, 3, dict(algo=1, cin=8), 0, 48),
    ("thin_proj_x16", 2, 8, 56, 56, 32, 8, 3, dict(algo=1, cin=8), 16, 0),
    ("thin_proj_x16_odd", 2, 5, 13, 11, 32, 8, 3, dict(algo=1, cin=8), 16, 0),
]

# full-size blocks of SlowFast-R50 8x8, batch 64 (Fast pathway res2 on 2-pixel groups, res3, res4)
BENCH_CASES = [
    ("bench_fast_s2_grouped", 64, 32, 56, 28, 64, 16, 3, {}, 0, 0),
    ("bench_fast_s2_plain", 64, 32, 56, 56, 32, 16, 3, {}, 0, 0),
    ("bench_fast_s3", 64, 32, 28, 28, 64, 16, 3, {}, 0, 0),
    ("bench_fast_s4", 64, 32, 14, 14, 128, 32, 3, {}, 0, 0),
    ("bench_thin_s2", 64, 32, 56, 56, 32, 8, 3, dict(algo=1), 0, 0),
    ("bench_thin_s3", 64, 32, 28, 28, 64, 16, 3, dict(algo=1), 0, 0),
    ("bench_thin_s2_slots4", 64, 32, 56, 56, 32, 8, 3, dict(algo=1, stages=4), 0, 0),
    ("bench_thin_s2_rows7", 64, 32, 56, 56, 32, 8, 3, dict(algo=1, walk_len=7), 0, 0),
    ("bench_thin_proj_s2", 64, 32, 56, 56, 32, 8, 3, dict(algo=1, cin=8), 0, 0),
    ("bench_thin_proj_s2_x16", 64, 32, 56, 56, 32, 8, 3, dict(algo=1, cin=8), 16, 0),
]


def make_case(case):
    name, n, t, h, w, c, d, kt, tune, xp, op = case
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()) % (2 ** 31))
    cin = tune.get("cin", 0) or c
    xp = xp or cin
    op = op or c
    xbuf = torch.randn((n, t, h, w, xp), generator=g).to(torch.bfloat16).to(dev)
    wa = (torch.randn((d, cin, kt, 1, 1), generator=g) / (cin * kt) ** 0.5).to(torch.bfloat16).to(dev)
    wb = (torch.randn((d, d, 1, 3, 3), generator=g) / (9 * d) ** 0.5).to(torch.bfloat16).to(dev)
    wc = (torch.randn((c, d, 1, 1, 1), generator=g) / d ** 0.5).to(torch.bfloat16).to(dev)
    aff = []
    for width in (d, d, c):
        aff.append((torch.rand(width, generator=g) + 0.5).to(dev))
        aff.append((torch.randn(width, generator=g) * 0.1).to(dev))
    outbuf = torch.full((n, t, h, w, op), 7.0, dtype=torch.bfloat16, device=dev)
    return xbuf, wa, wb, wc, aff, outbuf, xp, op


def reference(x, wa, wb, wc, aff, kt):
    sa, ba, sb, bb, sc, bc = aff
    v = lambda s: s.view(1, -1, 1, 1, 1)
    xc = x.float().permute(0, 4, 1, 2, 3)
    a = torch.relu(F.conv3d(xc, wa.float(), padding=(kt // 2, 0, 0)) * v(sa) + v(ba)).to(torch.bfloat16).float()
    b = torch.relu(F.conv3d(a, wb.float(), padding=(0, 1, 1)) * v(sb) + v(bb)).to(torch.bfloat16).float()
    y = torch.relu(F.conv3d(b, wc.float()) * v(sc) + v(bc) + xc)
    return y.permute(0, 2, 3, 4, 1)


def reference_proj(x, wa, wb, wcat, aff, kt, d):
    """Projection block as the kernel (and the engine's fused-shortcut conv) computes it: one accumulation over
    [b | x] with the folded bf16 weights, then the common scale and the summed bias."""
    sa, ba, sb, bb, s, bsum = aff
    v = lambda q: q.view(1, -1, 1, 1, 1)
    xc = x.float().permute(0, 4, 1, 2, 3)
    a = torch.relu(F.conv3d(xc, wa.float(), padding=(kt // 2, 0, 0)) * v(sa) + v(ba)).to(torch.bfloat16).float()
    b = torch.relu(F.conv3d(a, wb.float(), padding=(0, 1, 1)) * v(sb) + v(bb)).to(torch.bfloat16).float()
    wf = wcat.float()
    acc = F.conv3d(b, wf[:, :d, None, None, None]) + F.conv3d(xc, wf[:, d:, None, None, None])
    return torch.relu(acc * v(s) + v(bsum)).permute(0, 2, 3, 4, 1)


def run_case(case, bench=False):
    name, n, t, h, w, c, d, kt, tune, _, _ = case
    xbuf, wa, wb, wc, aff, outbuf, xp, op = make_case(case)
    cin = tune.get("cin", 0)
    wa_p = wa.permute(0, 2, 3, 4, 1).reshape(d, kt, cin or c).contiguous()
    wb_p = wb.permute(0, 2, 3, 4, 1).reshape(d, 9, d).contiguous()
    wc_p = wc.reshape(c, d).contiguous()
    if cin:
        # fold the two BatchNorm scales as ratios to the larger one (engine._conv_with_shortcut)
        g = torch.Generator(device="cpu").manual_seed(zlib.crc32((name + "proj").encode()) % (2 ** 31))
        w1 = (torch.randn((c, cin), generator=g) / cin ** 0.5).cuda()
        s_1 = (torch.rand(c, generator=g) + 0.5).cuda() * torch.where(torch.rand(c, generator=g) < 0.3, -1.0, 1.0).cuda()
        b_1 = (torch.randn(c, generator=g) * 0.1).cuda()
        s_c, b_c = aff[4], aff[5]
        use_c = s_c.abs() >= s_1.abs()
        s = torch.where(use_c, s_c, s_1)
        wc_p = torch.cat([wc_p.float() * (s_c / s)[:, None], w1 * (s_1 / s)[:, None]], dim=1).to(torch.bfloat16).contiguous()
        aff = aff[:4] + [s.contiguous(), (b_c + b_1).contiguous()]
    plan = BottleneckPlan(Act(xbuf, n, t, h, w, cin or c, xp), Act(outbuf, n, t, h, w, c, op), d, kt, wa_p, wb_p, wc_p, *aff,
                          **tune)
    if cin:
        reference_fn = lambda xx, *_: reference_proj(xx, wa, wb, wc_p, aff, kt, d)
    else:
        reference_fn = lambda xx, *_: reference(xx, wa, wb, wc, aff, kt)
    cx = cin or c
    info = {"plan": plan.info()}
    plan.run()
    torch.cuda.synchronize()
    if bench:
        for _ in range(3):
            plan.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            plan.run()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / iters
        px = n * t * h * w
        info["ms"] = ms
        info["hbm_gbs"] = px * (c + cx) * 2 / ms / 1e6
        info["tflops"] = 2.0 * px * (kt * c * d + 9 * d * d + d * c) / ms / 1e9
        # reference on a slice of the batch (memory)
        ns = 2
        ref = reference_fn(xbuf[:ns, ..., :cx])
        got = outbuf[:ns, ..., :c].float()
    else:
        ref = reference_fn(xbuf[..., :cx])
        got = outbuf[..., :c].float()
    err = (got - ref).abs()
    tol = 2e-2 + 1.6e-2 * ref.abs()
    bad = err > tol
    info.update(max_abs_err=float(err.max()), max_ref=float(ref.abs().max()), n_bad=int(bad.sum()), n=int(bad.numel()),
                finite=bool(torch.isfinite(got).all()))
    if info["n_bad"]:
        idx = bad.nonzero()
        info["first_bad"] = idx[:8].tolist()
        info["bad_per_t"] = bad.sum(dim=(0, 2, 3, 4)).tolist()
        info["bad_per_y"] = bad.sum(dim=(0, 1, 3, 4)).tolist()
        info["bad_per_x"] = bad.sum(dim=(0, 1, 2, 4)).tolist()
        i = idx[0].tolist()
        info["first_vals"] = {"got": float(got[tuple(i)]), "ref": float(ref[tuple(i)])}
    if op != c:
        info["slice_clean"] = bool((outbuf[..., c:] == 7.0).all())
        if not info["slice_clean"]:
            info["n_bad"] += 1
    info["ok"] = info["n_bad"] == 0 and info["finite"]
    return info


def child(args):
    report = json.load(open(args.out)) if os.path.exists(args.out) else {}
    report.setdefault("fused", {})

    def save():
        json.dump(report, open(args.out, "w"), indent=1)

    report["device"] = torch.cuda.get_device_name(0)
    try:
        cases = [(c, False) for c in CASES] + ([(c, True) for c in BENCH_CASES] if args.bench else [])
        if args.only:
            cases = [cb for cb in cases if args.only in cb[0][0]]
        for case, bench in cases:
            if case[0] in report["fused"]:
                continue
            report["fused"][case[0]] = {"ok": False, "crashed": True}
            save()
            try:
                report["fused"][case[0]] = run_case(case, bench)
            except Exception as e:
                if "CUDA" in str(e) or "cuda" in str(e) and "launch" in str(e):
                    raise
                from vidsitu_b200.lib import VsbError
                if isinstance(e, VsbError) and "plan_create" in str(e):
                    report["fused"][case[0]] = {"ok": False, "plan_error": str(e)[:400]}
                else:
                    raise
            save()
    except Exception as e:
        report.setdefault("exceptions", []).append(f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}")
        save()
        return 3
    report["complete"] = True
    save()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/check_fused.json")
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--only", default="", help="run only the cases whose name contains this string")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    if args.child:
        sys.exit(child(args))
    if os.path.exists(args.out):
        os.remove(args.out)
    for attempt in range(30):
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--out", args.out] + (["--bench"] if args.bench else [])
        if args.only:
            cmd += ["--only", args.only]
        try:
            rc = subprocess.run(cmd, timeout=300).returncode
        except subprocess.TimeoutExpired:
            rc = -9
        if rc == 0:
            break
        print(f"[check_fused] child exited with {rc}, restarting after the crashed case", flush=True)
        time.sleep(1)
    report = json.load(open(args.out)) if os.path.exists(args.out) else {}
    n_ok = n_bad = 0
    for k, v in report.get("fused", {}).items():
        ok = bool(v.get("ok"))
        n_ok += ok
        n_bad += not ok
        line = f"{'ok  ' if ok else 'FAIL'} {k}: " + json.dumps({kk: vv for kk, vv in v.items() if kk not in ('ok',)})[:700]
        print(line)
    print(f"[check_fused] ok={n_ok} bad={n_bad} complete={report.get('complete', False)}")
    for e in report.get("exceptions", [])[:3]:
        print("[exception]", e[:800])
    sys.exit(0 if n_bad == 0 and report.get("complete") else 1)


if __name__ == "__main__":
    main()
