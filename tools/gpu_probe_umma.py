"""Pins tcgen05.mma shared-memory descriptor semantics and issue rates on the GPU box.

    python tools/gpu_probe_umma.py  -> gpurun_out/umma_probe.json

1. semantics: D = A[shift rows, K slice] . B^T for descriptors whose start address is shifted by
   whole rows and by K slices inside a swizzled tile, with and without the base_offset field;
2. rate: cycles per M=128 x N x K=16 MMA for N = 16..256 (A tile streamed / reused).
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vidsitu_b200 import lib as L

lib = L.load()
dev = torch.device("cuda")
report = {"semantics": [], "rate": []}


def semantics(row_bytes, n, ksteps, shift_rows, a_koff, b_koff, base_mode):
    cols = row_bytes // 2
    a_rows = 256
    g = torch.Generator().manual_seed(row_bytes + n + shift_rows)
    a = torch.randint(-4, 5, (a_rows, cols), generator=g).to(torch.bfloat16)
    b = torch.randint(-4, 5, (n, cols), generator=g).to(torch.bfloat16)
    out = torch.full((128, n), float("nan"), dtype=torch.float32, device=dev)
    ad, bd = a.to(dev), b.to(dev)
    rc = lib.vsb_debug_umma_semantics(ad.data_ptr(), a_rows, bd.data_ptr(), n, row_bytes, ksteps,
                                      shift_rows * row_bytes + a_koff * 2, b_koff * 2, base_mode, out.data_ptr(), None)
    L.check(rc, "vsb_debug_umma_semantics")
    torch.cuda.synchronize()
    k = 16 * ksteps
    ref = a[shift_rows:shift_rows + 128, a_koff:a_koff + k].float() @ b[:, b_koff:b_koff + k].float().t()
    got = out.cpu()
    ok = bool(torch.equal(got, ref))
    bad_rows = sorted(set(torch.nonzero((got != ref).any(1)).flatten().tolist()))
    return ok, bad_rows[:12], len(bad_rows)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
for row_bytes in ((128, 64, 32) if which in ("all", "sem") else ()):
    cols = row_bytes // 2
    for base_mode in (0, 1):
        for shift in (0, 1, 2, 3, 5, 7, 8, 9, 57, 64, 115):
            for (ksteps, a_koff, b_koff) in ((cols // 16, 0, 0), (1, cols - 16, 0), (1, 0, cols - 16), (1, 16 % cols, 0)):
                if a_koff + 16 * ksteps > cols or b_koff + 16 * ksteps > cols:
                    continue
                try:
                    ok, bad, nbad = semantics(row_bytes, 32, ksteps, shift, a_koff, b_koff, base_mode)
                except Exception as e:  # noqa: BLE001
                    ok, bad, nbad = False, [str(e)[:200]], -1
                rec = {"row_bytes": row_bytes, "base_mode": base_mode, "shift_rows": shift, "ksteps": ksteps,
                       "a_koff": a_koff, "b_koff": b_koff, "ok": ok, "bad_rows": bad, "n_bad": nbad}
                report["semantics"].append(rec)
                print(rec, flush=True)

clk = torch.zeros(148 * 2, dtype=torch.int64, device=dev)
for grid, pad in (((148, 118), (296, 0)) if which in ("all", "rate") else ()):
    for same in (0, 1):
        for n in (16, 32, 64, 80, 96, 128, 160, 192, 208, 256):
            iters = 2000
            rc = lib.vsb_debug_umma_rate(n, iters, 4, same, grid, pad, clk.data_ptr(), None)
            L.check(rc, "vsb_debug_umma_rate")
            torch.cuda.synchronize()
            c = clk[:grid].float()
            rec = {"grid": grid, "a_reused": same, "n": n, "clk_per_mma_mean": float(c.mean()) / (iters * 4),
                   "clk_per_mma_max": float(c.max()) / (iters * 4)}
            report["rate"].append(rec)
            print(rec, flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(report, open(f"gpurun_out/umma_probe_{which}.json", "w"), indent=1)
