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
if "--mem" in sys.argv:
    for k, v in G.run_mem_checks().items():
        report("mem/" + k, v)
torch.cuda.synchronize()
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
