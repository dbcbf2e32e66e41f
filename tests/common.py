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


def torch_bf16_comparator(model, cfg, frames_cpu: torch.Tensor):
    """EXTERNAL comparator of the bf16 tolerance (test infrastructure): the reference's own modules - the oracle
    restatement, i.e. F.conv3d -> cuDNN, F.batch_norm, F.max_pool3d, einsum - run in bf16 channels_last_3d on the
    GPU with the same weights and clips; what `model.to(memory_format=torch.channels_last_3d).bfloat16()` gives a user
    of the reference.  Returns (pooled [N, D], logits [N, V]) as fp32 numpy arrays."""
    from oracle import sf_oracle as O
    dev = torch.device("cuda")
    sd = {k: (v.detach().to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5
              else v.detach().to(dev, torch.bfloat16)) for k, v in model.state_dict().items()}
    xs = [x.to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
          for x in O.clips_from_frames(frames_cpu, cfg.sf_mdl)]
    _, pooled, logits = O.sfbase_forward(sd, cfg.sf_mdl, xs)
    return pooled.float().cpu().numpy(), logits.float().cpu().numpy()


def synthetic_image(h: int, w: int, kind: str = "noisy", seed: int = 0):
    """uint8 [h, w, 3]: smooth colour gradients, optionally with noise (so that every AC coefficient is exercised)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 9.0 + yy / 17.0), 128 + 90 * np.cos(xx / 13.0 - yy / 7.0),
                     (xx * 3 + yy * 5) % 256], -1).astype(np.float64)
    if kind == "noisy":
        base = base + rng.normal(0, 40, base.shape)
    return np.clip(base, 0, 255).astype(np.uint8)


# (height, width, quality, PIL subsampling: 0 = 4:4:4, 1 = 4:2:2, 2 = 4:2:0, kind) - odd sizes, partial MCUs, the
# ffmpeg-like case (4:2:0, high quality, 16:9)
JPEG_CASES = [(64, 80, 95, 2, "smooth"), (77, 53, 90, 2, "noisy"), (48, 64, 75, 0, "noisy"), (50, 70, 85, 1, "smooth"),
              (120, 160, 98, 2, "noisy"), (33, 47, 60, 2, "noisy"), (16, 16, 100, 2, "noisy"), (90, 122, 92, 1, "noisy"),
              (360, 640, 96, 2, "noisy"), (240, 426, 93, 2, "smooth")]


def jpeg_bytes(h: int, w: int, quality: int, subsampling: int, kind: str, seed: int = 0, gray: bool = False) -> bytes:
    import io
    from PIL import Image
    a = synthetic_image(h, w, kind, seed)
    buf = io.BytesIO()
    if gray:
        Image.fromarray(a[..., 0]).save(buf, "JPEG", quality=quality)
    else:
        Image.fromarray(a).save(buf, "JPEG", quality=quality, subsampling=subsampling)
    return buf.getvalue()
