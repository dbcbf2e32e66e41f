"""On-disk feature contract of the hot path.

The reference writes one file per video, `{out_tdir}/{vseg_name}_feats.npy`, holding the
fp32 `[5, D]` pooled event features (event-major, D = [slow | fast]) -- the tail of
`FeatExtract.forward_all` (vidsitu_code/feat_extractor.py:98-111) -- and the SRL / event-relation
models read it back with `get_frm_feats_all` (vidsitu_code/dat_loader.py:503-511), choosing the
feature width from the directory name (`get_head_dim`, vidsitu_code/mdl_sf_base.py:751-760).

`FeatureWriter` is that tail without the per-batch stall: pooled features come back through a
ring of pinned host buffers (asynchronous D2H on the producing stream, one CUDA event per slot)
and a background thread does the `np.save` calls while the next batch is in flight.  The bytes
on disk are exactly what `np.save(path, out_np_one)` of the reference produces: C-contiguous
little-endian float32 `[5, D]`, .npy format version 1.0.
"""
from __future__ import annotations

import os
import queue
import threading
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np
import torch

EVENTS_PER_VIDEO = 5  # Ev1..Ev5, dat_loader.py:454-472


def feats_path(out_dir, vseg_name: str) -> Path:
    """`{out_tdir}/{vseg_name}_feats.npy` (feat_extractor.py:109, dat_loader.py:505-507)."""
    return Path(out_dir) / f"{vseg_name}_feats.npy"


def get_head_dim(frm_feats_dir: str) -> int:
    """Feature width implied by the directory name (mdl_sf_base.py:751-760)."""
    d = str(frm_feats_dir)
    if "i3d" in d:
        return 2048
    if "slow_fast" in d or "sfast" in d:
        return 2304
    raise NotImplementedError(f"no feature width rule for {d!r}")


def write_video_feats(out_dir, vseg_name: str, feats_5xd: np.ndarray) -> Path:
    """One video's features, exactly as `np.save(out_np_name, out_np_one)` (feat_extractor.py:108-110)."""
    a = np.ascontiguousarray(feats_5xd, dtype=np.float32)
    if a.ndim != 2 or a.shape[0] != EVENTS_PER_VIDEO:
        raise ValueError(f"expected [5, D] features, got {a.shape}")
    p = feats_path(out_dir, vseg_name)
    np.save(p, a)
    return p


def read_frm_feats(frm_feats_dir, vseg_name: str) -> torch.Tensor:
    """`get_frm_feats_all` (dat_loader.py:503-511): fp32 [5, D] tensor; asserts the event count and,
    when the directory name implies one, the width."""
    p = feats_path(frm_feats_dir, vseg_name)
    if not p.exists():
        raise AssertionError(f"{p} doesn't exist")   # read_file_with_assertion, dat_loader.py:40-44
    t = torch.from_numpy(np.load(p)).float()
    assert t.size(0) == EVENTS_PER_VIDEO
    try:
        want = get_head_dim(str(frm_feats_dir))
    except NotImplementedError:
        want = None
    if want is not None and t.size(1) != want:
        raise AssertionError(f"{p}: feature width {t.size(1)} != {want} implied by the directory name")
    return t


class FeatureWriter:
    """Asynchronous `[B, 5, D]` -> one .npy per video.

        w = FeatureWriter(cfg.ds.vsitu.vsitu_frm_feats, mdl_name)     # feat_extractor.py:86-88
        w.put(feats_device_or_host [5*B, D] or [B, 5, D], vseg_names)
        ...
        w.close()

    `put` with a CUDA tensor enqueues a non-blocking copy into a pinned slot on the current stream
    and returns at once; a slot is reused only after its files are on disk."""

    def __init__(self, frm_feats_root, mdl_name: Optional[str] = None, slots: int = 4, max_rows: int = 0):
        out = Path(frm_feats_root) / mdl_name if mdl_name else Path(frm_feats_root)
        out.mkdir(exist_ok=True, parents=True)
        self.out_dir = out
        self._slots = max(2, int(slots))
        self._bufs: List[Optional[torch.Tensor]] = [None] * self._slots
        self._free: "queue.Queue[int]" = queue.Queue()
        for i in range(self._slots):
            self._free.put(i)
        self._work: "queue.Queue" = queue.Queue()
        self._err: Optional[BaseException] = None
        self.files_written = 0
        self._pin = torch.cuda.is_available()
        self._max_rows = max_rows
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def _loop(self) -> None:
        while True:
            item = self._work.get()
            if item is None:
                return
            slot, rows, d, names, ev = item
            try:
                if ev is not None:
                    ev.synchronize()
                a = self._bufs[slot][: rows * d].view(rows // EVENTS_PER_VIDEO, EVENTS_PER_VIDEO, d).numpy()
                for v, name in enumerate(names):
                    write_video_feats(self.out_dir, name, a[v])
                    self.files_written += 1
            except BaseException as e:  # surfaced by the next put() / close()
                self._err = e
            finally:
                self._free.put(slot)

    def put(self, feats: torch.Tensor, vseg_names: Sequence[str]) -> None:
        if self._err is not None:
            raise self._err
        if feats.dim() == 3:
            if feats.shape[1] != EVENTS_PER_VIDEO:
                raise ValueError("expected [B, 5, D]")
            feats = feats.reshape(-1, feats.shape[-1])
        rows, d = feats.shape
        if rows != EVENTS_PER_VIDEO * len(vseg_names):
            raise ValueError(f"{rows} feature rows for {len(vseg_names)} videos (5 events each)")
        if feats.dtype != torch.float32:
            raise ValueError("features are written as float32, as the reference does")
        slot = self._free.get()
        need = max(rows * d, self._max_rows * d)
        buf = self._bufs[slot]
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.float32)
            if self._pin:
                buf = buf.pin_memory()
            self._bufs[slot] = buf
        dst = buf[: rows * d].view(rows, d)
        ev = None
        if feats.is_cuda:
            dst.copy_(feats, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(feats.device))
        else:
            dst.copy_(feats)
        self._work.put((slot, rows, d, list(vseg_names), ev))

    def close(self) -> None:
        self._work.put(None)
        self._thr.join()
        if self._err is not None:
            raise self._err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
