"""End-to-end feature dump from JPEG files (tools/extract_features.py) with the three readers: PIL in DataLoader
workers (the reference's reader), --gpu-decode hybrid, --gpu-decode device.  Synthetic dataset: V videos, the 150
frames the event windows need, 640x360 4:2:0 JPEGs."""
import json, os, sys, tempfile, time
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))


def make_video(args):
    tdir, v, needed, quality = args
    from PIL import Image
    from common import synthetic_image
    d = os.path.join(tdir, v)
    os.makedirs(d, exist_ok=True)
    total = 0
    for i in needed:
        p = os.path.join(d, f"{v}_{i + 1:06d}.jpg")
        Image.fromarray(synthetic_image(360, 640, "smooth" if i % 4 else "noisy", seed=hash(v) % 1000 + i)).save(
            p, quality=quality, subsampling=2)
        total += os.path.getsize(p)
    return total


def main():
    """One warm model, the three loaders iterated over the same V videos; the clock covers the whole iteration (the
    DataLoader's worker start-up and prefetching included - its workers decode ahead of the first delivered batch, so
    any later starting point would hide most of their work)."""
    import torch
    from torch.utils.data import DataLoader
    from common import build_model
    from vidsitu_b200 import frames_io as F
    V = int(os.environ.get("VIDEOS", "96"))
    B = int(os.environ.get("VPB", "8"))
    needed = F.needed_frames(32, 2)
    res = {"videos": V, "frames_per_video": len(needed), "frame": "640x360 4:2:0 quality 90", "host_cpus": os.cpu_count(),
           "videos_per_batch": B}
    model, cfg, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=224, micro_batch=5 * B)
    model = model.cuda()
    with tempfile.TemporaryDirectory() as tmp:
        names = [f"v_{k:03d}_seg_0" for k in range(V)]
        with ProcessPoolExecutor(16) as pool:
            sizes = list(pool.map(make_video, [(os.path.join(tmp, "frames"), v, needed, 90) for v in names]))
        res["mean_jpeg_kb"] = round(sum(sizes) / (V * len(needed)) / 1024, 1)
        tdir = os.path.join(tmp, "frames")

        def loaders():
            ds = F.VideoFrames(tdir, names, 32, 2, 30, size=224)
            yield "pil_16_workers", DataLoader(ds, batch_size=B, shuffle=False, num_workers=16, collate_fn=F.collate_videos,
                                               pin_memory=True)
            yield "gpu_hybrid_16_threads", F.DeviceVideoLoader(tdir, names, 32, 2, 30, videos_per_batch=B, workers=16, mode="hybrid")
            yield "gpu_device_1_thread", F.DeviceVideoLoader(tdir, names, 32, 2, 30, videos_per_batch=B, workers=4, mode="device")
        feats = {}
        model.extract_video_features(torch.zeros((B, 300, 224, 224, 3), dtype=torch.uint8, device="cuda"))   # engine + graph
        torch.cuda.synchronize()
        for mode, dl in loaders():
            out = []
            t0 = time.time()
            for frames, idxs in dl:
                f = model.extract_video_features(frames.cuda(non_blocking=True))
                torch.cuda.synchronize()
                out.append(f.cpu())
            dt = time.time() - t0
            res[mode + "_clips_per_s"] = round(5 * V / dt, 1)
            res[mode + "_frames_per_s"] = round(len(needed) * V / dt, 0)
            res[mode + "_s"] = round(dt, 2)
            feats[mode] = torch.cat(out)
        res["identical_features"] = all(torch.equal(feats["pil_16_workers"], v) for v in feats.values())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
