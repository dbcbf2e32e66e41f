#!/usr/bin/env python
"""Export the forward of a VidSitu SlowFast / I3D model as a clip program file (C ABI v7, include/vidsitu_b200.h):
the artifact a non-Python host loads with vsb_program_load and runs with vsb_program_run (examples/run_program.c).

    python tools/export_program.py --sf-mdl-name slow_fast_nl_r50_8x8 --clips 64 --out sf50_n64.vsbprog \
        [--mdl-resume-path ckpt.pth | --mdl-resume-path SLOWFAST_8x8_R50.pkl --is-cu] [--crop 224] [--precision bf16]

The program is specialised to the batch size (`--clips`): buffers, TMA descriptors and grids are planned for it.
Needs a B200 (plans are validated against the device at build time)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf-mdl-name", default="slow_fast_nl_r50_8x8")
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--crop", type=int, default=224)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--mdl-resume-path", default="")
    ap.add_argument("--is-cu", action="store_true")
    ap.add_argument("--seed", type=int, default=0, help="random-init seed when no checkpoint is given")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    import torch
    from common import build_model
    from vidsitu_b200 import checkpoint

    mdl, cfg, _ = build_model(args.sf_mdl_name, seed=args.seed, crop=args.crop, precision=args.precision)
    if args.mdl_resume_path:
        if args.is_cu:
            print(checkpoint.load_caffe2_checkpoint(args.mdl_resume_path, mdl.sf_mdl))
        else:
            checkpoint.load_vidsitu_checkpoint(args.mdl_resume_path, mdl)
    mdl = mdl.cuda()
    eng = mdl._engine(args.clips, torch.device("cuda"))
    prog = eng.export_program(args.out)
    print(f"{args.out}: {prog.num_launches} launches, {prog.device_bytes / 2**20:.1f} MiB of device memory, "
          f"{os.path.getsize(args.out) / 2**20:.1f} MiB on disk")


if __name__ == "__main__":
    main()
