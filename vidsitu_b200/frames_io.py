"""Frame ingest of the offline feature dump (host side).

The reference stores every 10-second video as 300 JPEGs `{video_frms_tdir}/{vseg}/{vseg}_{i:06d}.jpg`
(i = 1..300, 30 fps; `VsituDS.get_frms_all` vidsitu_code/dat_loader.py:454-458) and, per event, opens
the NUM_FRAMES frames of the window with PIL, converts to RGB and resizes to 224 x 224
(`read_img` :183-191), then normalises and packs them on the CPU (:474-486) - 160 decodes and
5 x 24 MB of fp32 per video.  Here a video is decoded ONCE into a uint8 `[300, H, W, 3]` tensor, only
the frames some event window selects (`events.event_frame_indices`; at most 150 distinct ones for
SlowFast 8x8), and everything after the decode - windowing, normalisation, pathway packing - runs in
the pack kernel on the device-resident video (`SFBase.extract_video_features`).

JPEG decoding itself stays on the host (PIL, exactly the reference's three calls, in DataLoader
worker processes): it is outside the hot path's scope.
"""
from __future__ import annotations

import json
import pickle
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

from .events import VIDEO_FRAMES, event_frame_indices


def frame_paths(video_frms_tdir, vseg: str, num_frames: int = VIDEO_FRAMES) -> List[Path]:
    """`{tdir}/{vseg}/{vseg}_{ix:06d}.jpg` for ix = 1..300 (dat_loader.py:455-458): list index = frame index."""
    d = Path(video_frms_tdir) / vseg
    return [d / f"{vseg}_{ix:06d}.jpg" for ix in range(1, num_frames + 1)]


def read_img(img_fpath, size: int = 224) -> np.ndarray:
    """H x W x 3 uint8, as `VsituDS.read_img` (dat_loader.py:183-191): open, RGB, `resize((224, 224))` with
    PIL's default filter."""
    from PIL import Image
    img = Image.open(img_fpath).convert("RGB")
    img = img.resize((size, size))
    return np.array(img)


def needed_frames(num_frames: int, sampling_rate: int, fps: int = 30, total: int = VIDEO_FRAMES) -> List[int]:
    """Distinct frame indices any of the five event windows selects."""
    return sorted({i for win in event_frame_indices(num_frames, sampling_rate, fps, total) for i in win})


def load_video(video_frms_tdir, vseg: str, needed: Sequence[int], size: int = 224,
               total: int = VIDEO_FRAMES) -> torch.Tensor:
    """uint8 [total, size, size, 3]; frames outside `needed` stay zero (no window reads them)."""
    paths = frame_paths(video_frms_tdir, vseg, total)
    out = np.zeros((total, size, size, 3), dtype=np.uint8)
    for i in needed:
        if not paths[i].exists():
            raise AssertionError(f"{paths[i]} doesn't exist")     # read_file_with_assertion semantics
        out[i] = read_img(paths[i], size)
    return torch.from_numpy(out)


def read_vseg_list(path) -> List[str]:
    """A split file of the reference (`cfg.ds.vsitu.split_files_lb[split]`: JSON list of vseg names; pickle and
    one-name-per-line text are accepted too)."""
    p = Path(path)
    if not p.exists():
        raise AssertionError(f"{p} doesn't exist")
    if p.suffix == ".json":
        return list(json.load(open(p)))
    if p.suffix in (".pkl", ".pickle"):
        return list(pickle.load(open(p, "rb")))
    return [ln.strip() for ln in open(p) if ln.strip()]


class VideoFrames(Dataset):
    """One item = one video: (`frames` uint8 [300, size, size, 3], index).  The analogue of `VsituDS_All`
    (vidsitu_code/feat_extractor.py:20-74) for the whole-video path: DataLoader workers do the JPEG decodes,
    the main process only moves uint8 tensors."""

    def __init__(self, video_frms_tdir, vseg_lst: Sequence[str], num_frames: int, sampling_rate: int,
                 fps: int = 30, size: int = 224, total: int = VIDEO_FRAMES):
        self.tdir = Path(video_frms_tdir)
        self.vseg_lst = list(vseg_lst)
        self.size, self.total = int(size), int(total)
        self.needed = needed_frames(num_frames, sampling_rate, fps, total)

    def __len__(self) -> int:
        return len(self.vseg_lst)

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, int]:
        return load_video(self.tdir, self.vseg_lst[idx], self.needed, self.size, self.total), idx


def collate_videos(batch):
    frames = torch.stack([b[0] for b in batch])
    return frames, [b[1] for b in batch]


# ------------------------------------------------------------------------------------ GPU decode
class DeviceVideoLoader:
    """Whole videos decoded ON THE GPU: the counterpart of DataLoader(VideoFrames) for `--gpu-decode`.

    mode "device" (default): `jpeg.JpegBatchDecoder` - the JPEG files of a loader batch (videos_per_batch x ~150 needed
    frames) are read by `workers` I/O threads and decoded in ONE library call, Huffman segments included (one warp
    per frame), straight into the videos' uint8 `[total, size, size, 3]` device tensors.
    mode "hybrid": `workers` host threads each own a `jpeg.JpegDecoder` and a CUDA stream and Huffman-decode their
    videos' files on the host (in the library, outside the GIL); the pixel stages run on the device.
    Either way no pixel crosses PCIe and the frames are bit-identical to the reference's PIL reader.  Files the GPU
    path refuses (progressive, CMYK, ...) go through that reader (`read_img`) and are counted in `host_fallbacks`.
    Iterating yields `(frames [B, total, size, size, 3] cuda uint8, [video indices])`, ready on the CURRENT stream."""

    def __init__(self, video_frms_tdir, vseg_lst: Sequence[str], num_frames: int, sampling_rate: int, fps: int = 30,
                 size: int = 224, total: int = VIDEO_FRAMES, videos_per_batch: int = 8, workers: int = 4,
                 device=None, max_width: int = 1920, max_height: int = 1088, mode: str = "device"):
        import threading
        if mode not in ("device", "hybrid"):
            raise ValueError("mode must be 'device' or 'hybrid'")
        self.tdir = Path(video_frms_tdir)
        self.vseg_lst = list(vseg_lst)
        self.size, self.total = int(size), int(total)
        self.needed = needed_frames(num_frames, sampling_rate, fps, total)
        self.videos_per_batch = int(videos_per_batch)
        self.workers = max(1, int(workers))
        self.device = torch.device(device if device is not None else "cuda")
        self.max_wh = (int(max_width), int(max_height))
        self.mode = mode
        self.host_fallbacks = 0
        self._tls = threading.local()
        self._lock = threading.Lock()
        self._batch = None

    def __len__(self) -> int:
        return (len(self.vseg_lst) + self.videos_per_batch - 1) // self.videos_per_batch

    # ---- hybrid: one decoder + stream per host thread
    def _state(self):
        from .jpeg import JpegDecoder
        st = getattr(self._tls, "st", None)
        if st is None:
            st = (JpegDecoder(self.max_wh[0], self.max_wh[1], self.device), torch.cuda.Stream(self.device))
            self._tls.st = st
        return st

    def _paths(self, vseg: str):
        paths = frame_paths(self.tdir, vseg, self.total)
        for i in self.needed:
            if not paths[i].exists():
                raise AssertionError(f"{paths[i]} doesn't exist")     # read_file_with_assertion semantics
        return paths

    def _host_frame(self, path, dst: torch.Tensor) -> None:
        dst.copy_(torch.from_numpy(read_img(path, self.size)))
        with self._lock:
            self.host_fallbacks += 1

    def _decode_video(self, vseg: str, out: torch.Tensor) -> "torch.cuda.Event":
        from .lib import VsbError
        dec, stream = self._state()
        paths = self._paths(vseg)
        with torch.cuda.device(self.device), torch.cuda.stream(stream):
            for i in self.needed:
                try:
                    dec.decode_resize(paths[i].read_bytes(), out[i])
                except VsbError:
                    self._host_frame(paths[i], out[i])
            ev = torch.cuda.Event()
            ev.record(stream)
        return ev

    # ---- device: the whole loader batch in one call
    def _read_batch(self, pool, idxs):
        """Start reading the JPEG files of a loader batch on the I/O threads: (jobs, futures of the file bytes)."""
        jobs = [(k, i, p[i]) for k, v in enumerate(idxs) for p in (self._paths(self.vseg_lst[v]),) for i in self.needed]
        return jobs, [pool.submit(j[2].read_bytes) for j in jobs]

    def _decode_batch(self, jobs, futs, frames: torch.Tensor) -> None:
        from .jpeg import JpegBatchDecoder
        if self._batch is None:
            self._batch = JpegBatchDecoder(self.device)
        datas = [f.result() for f in futs]
        ok = self._batch.decode_resize(datas, [frames[k, i] for k, i, _ in jobs])
        for good, (k, i, path) in zip(ok, jobs):
            if not good:
                self._host_frame(path, frames[k, i])

    def __iter__(self):
        from concurrent.futures import ThreadPoolExecutor
        batches = [list(range(b0, min(b0 + self.videos_per_batch, len(self.vseg_lst))))
                   for b0 in range(0, len(self.vseg_lst), self.videos_per_batch)]
        with ThreadPoolExecutor(max_workers=self.workers) as pool:
            if self.mode == "device":
                # the files of batch k + 1 are read while batch k is decoded (and while the caller's CNN runs on it)
                reads = self._read_batch(pool, batches[0]) if batches else None
                for bi, idxs in enumerate(batches):
                    jobs, futs = reads
                    reads = self._read_batch(pool, batches[bi + 1]) if bi + 1 < len(batches) else None
                    frames = torch.zeros((len(idxs), self.total, self.size, self.size, 3), dtype=torch.uint8,
                                         device=self.device)
                    with torch.cuda.device(self.device):
                        self._decode_batch(jobs, futs, frames)       # completes before returning
                    yield frames, idxs
                return
            pending = None
            for idxs in batches:
                frames = torch.zeros((len(idxs), self.total, self.size, self.size, 3), dtype=torch.uint8,
                                     device=self.device)
                torch.cuda.current_stream(self.device).synchronize()     # the zero fill precedes the workers' writes
                futs = [pool.submit(self._decode_video, self.vseg_lst[i], frames[k]) for k, i in enumerate(idxs)]
                if pending is not None:
                    yield self._finish(*pending)
                pending = (frames, idxs, futs)
            if pending is not None:
                yield self._finish(*pending)

    def _finish(self, frames, idxs, futs):
        for f in futs:
            torch.cuda.current_stream(self.device).wait_event(f.result())
        return frames, idxs
