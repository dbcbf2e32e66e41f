"""Shared helpers of the test-suite, the golden generator and bench.py: seeded synthetic
weights / clips that are reproducible on any machine (CPU torch RNG only)."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vidsitu_b200.config import make_cfg, make_comm  # noqa: E402
from vidsitu_b200.model import randomize_bn_  # noqa: E402

NUM_VERBS = 1560  # synthetic verb vocabulary size


def build_model(sf_mdl_name: str, seed: int = 0, randomize_bn: bool = True, precision: str = "bf16", crop: int = 224,
                sf_overrides: dict = None, **kw):
    """vidsitu_b200.SFBase with seeded random-init weights (torch.manual_seed(seed)) and, by
    default, seeded non-trivial BatchNorm statistics (SURVEY.md section 7 hard part 1)."""
    from vidsitu_b200.sf_base import SFBase

    cfg = make_cfg(sf_mdl_name, **(sf_overrides or {}))
    cfg.sf_mdl.DATA.CROP_SIZE = crop
    comm = make_comm(cfg.sf_mdl, NUM_VERBS)
    torch.manual_seed(seed)
    m = SFBase(cfg, comm, precision=precision, **kw)
    if randomize_bn:
        randomize_bn_(m, seed + 1)
    m.eval()
    return m, cfg, comm


def synthetic_frames(n: int, t: int, crop: int = 224, seed: int = 1234) -> torch.Tensor:
    """uint8 [n, t, crop, crop, 3] uniform bytes (SURVEY.md section 8d generator)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, t, crop, crop, 3), dtype=torch.uint8, generator=g)
