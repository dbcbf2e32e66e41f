#!/usr/bin/env python
"""Offline feature dump with the B200-native forward: the counterpart of
`python vidsitu_code/feat_extractor.py --mdl_resume_path ... --mdl_name_used ...` (data/DATA_PREP.md:137-157,
vidsitu_code/feat_extractor.py:120-175) for one split.

    python tools/extract_features.py --frames-dir <video_frms_tdir> --split-file <vseg list .json> \
        --out-dir <vsitu_frm_feats> --mdl-name-used sfast_kpret --sf-mdl-name slow_fast_nl_r50_8x8 \
        [--mdl-resume-path ckpt.pth | --mdl-resume-path SLOWFAST_8x8_R50.pkl --is-cu] [--videos-per-batch 8]

Per video: the JPEGs the five event windows select are decoded once by DataLoader workers (PIL, the reference's
`read_img`), the uint8 video goes to the GPU, `SFBase.extract_video_features` cuts the windows / normalises /
packs / runs the CNN, and `FeatureWriter` saves `{out_dir}/{mdl_name_used}/{vseg}_feats.npy` (fp32 [5, D]) - the
files `get_frm_feats_all` (vidsitu_code/dat_loader.py:503-511) reads.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from vidsitu_b200 import checkpoint  # noqa: E402
from vidsitu_b200.config import make_cfg, make_comm  # noqa: E402
from vidsitu_b200.feat_io import FeatureWriter  # noqa: E402
from vidsitu_b200.frames_io import DeviceVideoLoader, VideoFrames, collate_videos, read_vseg_list  # noqa: E402
from vidsitu_b200.sf_base import SFBase  # noqa: E402


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--frames-dir", required=True, help="cfg.ds.vsitu.video_frms_tdir")
    ap.add_argument("--split-file", required=True, help="list of vseg names (cfg.ds.vsitu.split_files_lb[split])")
    ap.add_argument("--out-dir", required=True, help="cfg.ds.vsitu.vsitu_frm_feats")
    ap.add_argument("--mdl-name-used", required=True, help="sub-directory of the features (feat_extractor.py:86-88)")
    ap.add_argument("--sf-mdl-name", default="slow_fast_nl_r50_8x8", help="mdl.sf_mdl_name (extended_config.py:14-20)")
    ap.add_argument("--mdl-resume-path", default="", help="VidSitu .pth, or a PySlowFast Caffe2 .pkl with --is-cu")
    ap.add_argument("--is-cu", action="store_true", help="the checkpoint is a Caffe2 pickle (feat_extractor.py:155-161)")
    ap.add_argument("--num-verbs", type=int, default=1560, help="len(comm.vb_id_vocab) when no checkpoint fixes it")
    ap.add_argument("--videos-per-batch", type=int, default=8)
    ap.add_argument("--workers", type=int, default=4, help="cfg.train.nwv")
    ap.add_argument("--crop", type=int, default=224)
    ap.add_argument("--gpu-decode", nargs="?", const="device", default="", choices=["device", "hybrid"],
                    help="decode + resize the JPEGs on the GPU, bit-identical to the PIL reader, instead of PIL in "
                         "DataLoader workers: 'device' (default) = Huffman segments on the GPU too, one library call per "
                         "batch; 'hybrid' = --workers host threads do the Huffman part")
    args = ap.parse_args(argv)

    cfg = make_cfg(args.sf_mdl_name)
    cfg.sf_mdl.DATA.CROP_SIZE = args.crop
    comm = make_comm(cfg.sf_mdl, args.num_verbs)
    mdl = SFBase(cfg, comm, micro_batch=5 * args.videos_per_batch)
    if args.mdl_resume_path:
        if args.is_cu:
            print("Using Caffe2 checkpoint")
            rep = checkpoint.load_caffe2_checkpoint(args.mdl_resume_path, mdl.sf_mdl)
            print(f"  loaded {len(rep['loaded'])} tensors; shape-mismatched blobs (not loaded): {rep['mismatched']}; "
                  f"blobs without a model tensor: {len(rep['skipped'])}; model tensors not in the checkpoint: "
                  f"{[k for k in rep['missing'] if not k.endswith('num_batches_tracked')]}")
        else:
            checkpoint.load_vidsitu_checkpoint(args.mdl_resume_path, mdl)
    mdl = mdl.to(torch.device("cuda")).eval()

    vsegs = read_vseg_list(args.split_file)
    d = cfg.sf_mdl.DATA
    if args.gpu_decode:
        dl = DeviceVideoLoader(args.frames_dir, vsegs, d.NUM_FRAMES, d.SAMPLING_RATE, d.TARGET_FPS, size=args.crop,
                               videos_per_batch=args.videos_per_batch, workers=max(1, args.workers), mode=args.gpu_decode)
    else:
        ds = VideoFrames(args.frames_dir, vsegs, d.NUM_FRAMES, d.SAMPLING_RATE, d.TARGET_FPS, size=args.crop)
        dl = DataLoader(ds, batch_size=args.videos_per_batch, shuffle=False, num_workers=args.workers,
                        collate_fn=collate_videos, pin_memory=True, drop_last=False)
    t0 = time.time()
    done = 0
    with FeatureWriter(args.out_dir, args.mdl_name_used) as writer:
        for frames, idxs in dl:
            feats = mdl.extract_video_features(frames.cuda(non_blocking=True))      # [B, 5, D] fp32
            if args.gpu_decode:
                torch.cuda.current_stream().synchronize()    # the batch's frame tensor is dropped on the next iteration
            writer.put(feats, [vsegs[i] for i in idxs])
            done += len(idxs)
    dt = time.time() - t0
    print(f"{done} videos ({5 * done} event clips) -> {writer.out_dir} in {dt:.1f} s ({5 * done / max(dt, 1e-9):.1f} clips/s "
          f"including JPEG decode" + (f" on the GPU, {dl.host_fallbacks} files through PIL)" if args.gpu_decode else ")"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
