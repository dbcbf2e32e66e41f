"""Host-facing feature extraction (the loop of feat_extractor.py:90-112 without the JPEG decode):
uint8 frames in pinned host memory -> pooled [N, D] fp32 features back in pinned host memory.

Double-buffered: the H2D copy of batch k+1 (copy stream) overlaps the forward of batch k
(compute stream); the D2H of batch k's features follows its forward on the compute stream.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .lib import VsbError


class HostPipeline:
    def __init__(self, model, batch: int, device=None, want_logits: bool = False, slots: int = 2):
        if not torch.cuda.is_available():
            raise VsbError("HostPipeline needs a CUDA device")
        self.model = model
        self.batch = int(batch)
        self.device = torch.device(device if device is not None else "cuda")
        spec = model.spec
        self.eng = model._engine(self.batch, self.device)
        self.want_logits = want_logits
        shape = (self.batch, spec.num_frames, spec.crop, spec.crop, 3)
        self.stage = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        d = self.eng.feats.shape[1]
        self.out_feats = [torch.empty((self.batch, d), dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.out_logits = ([torch.empty(tuple(self.eng.logits.shape), dtype=torch.float32).pin_memory()
                            for _ in range(slots)] if want_logits else None)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.h2d_done = [torch.cuda.Event() for _ in range(slots)]
        self.slot_free = [torch.cuda.Event() for _ in range(slots)]
        self.done = [torch.cuda.Event() for _ in range(slots)]
        self._k = 0
        self.eng_slots = len(self.eng.input_sets)
        self.in_free = [torch.cuda.Event() for _ in range(self.eng_slots)]
        self.h2d_bytes = self.stage[0].numel()
        self.d2h_bytes = self.out_feats[0].numel() * 4 + (self.out_logits[0].numel() * 4 if want_logits else 0)
        with torch.cuda.stream(self.compute_stream):
            self.eng.capture()
        for e in self.slot_free + self.in_free:
            e.record(self.compute_stream)

    def submit(self, frames_host: torch.Tensor) -> int:
        """Enqueue one batch (pinned uint8 [batch, T, H, W, 3]); returns its slot."""
        if not frames_host.is_pinned():
            raise VsbError("frames must live in pinned host memory")
        s = self._k % len(self.stage)
        self._k += 1
        es = s % self.eng_slots                                   # input slot of the engine
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.slot_free[s])      # the pack kernels of the previous user are done
            self.stage[s].copy_(frames_host, non_blocking=True)
            if self.eng_slots > 1:
                # pack on the copy stream into the engine's other input slot: it overlaps the previous batch's
                # trunk, which still reads its own slot (freed when that trunk's stems are done = in_free)
                self.copy_stream.wait_event(self.in_free[es])
                self.eng.load_frames(self.stage[s], es)
                self.slot_free[s].record(self.copy_stream)
            self.h2d_done[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[s])
            if self.eng_slots == 1:
                self.eng.load_frames(self.stage[s])
                self.slot_free[s].record(self.compute_stream)
            self.eng.replay(es)
            self.in_free[es].record(self.compute_stream)
            self.out_feats[s].copy_(self.eng.feats, non_blocking=True)
            if self.want_logits:
                self.out_logits[s].copy_(self.eng.logits, non_blocking=True)
            self.done[s].record(self.compute_stream)
        return s

    def result(self, slot: int):
        self.done[slot].synchronize()
        return (self.out_feats[slot], self.out_logits[slot]) if self.want_logits else self.out_feats[slot]

    def flush(self) -> None:
        self.compute_stream.synchronize()
        self.copy_stream.synchronize()
