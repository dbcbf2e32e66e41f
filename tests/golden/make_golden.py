"""Generates tests/golden/*.npz by running the UNMODIFIED reference (vidsitu_code.mdl_sf_base.SFBase
from /root/reference) on seeded synthetic clips and weights.  Runs only in the build container
(the reference tree does not exist on the GPU box); the fixtures it writes are committed.

    python tests/golden/make_golden.py [case ...]

Missing third-party imports of the reference are replaced by in-memory stub modules
 -- no reference file is modified or copied.
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

# (case name, sf_mdl_name, clips, crop, seed)
CASES = {
    "sf50_n5_224": ("slow_fast_nl_r50_8x8", 5, 224, 0),       # BASELINE.json config 1
    "sf50_n2_64": ("slow_fast_nl_r50_8x8", 2, 64, 10),
    "sf50_rawinit_n2_64": ("slow_fast_nl_r50_8x8", 2, 64, 11),  # stock init (zero final-BN gamma), untouched BN
    "i3d_n2_224": ("i3d_r50_8x8", 2, 224, 1),
    "i3d_nln_n2_224": ("i3d_r50_nl_8x8", 2, 224, 2),
    "i3d_nln_n2_64": ("i3d_r50_nl_8x8", 2, 64, 12),
    "slow_n2_64": ("slow_nl_r50_8x8", 2, 64, 3),
    "c2d_n2_64": ("c2d_r50_8x8", 2, 64, 4),
    "sf101_n1_224": ("slow_fast_r101_16x8", 1, 224, 5),
    "sf101_n2_64": ("slow_fast_r101_16x8", 2, 64, 13),
    # BN.NORM_TYPE = sub_batchnorm (SubBatchNorm3d: `<bn>.bn.*` aggregated + `<bn>.split_bn.*` per-split statistics)
    "sf50_subbn_n2_64": ("slow_fast_nl_r50_8x8", 2, 64, 14, {"BN": {"NORM_TYPE": "sub_batchnorm", "NUM_SPLITS": 2}}),
}


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Registry:
        def __init__(self, name):
            self._d = {}

        def register(self, obj=None):
            def deco(o):
                self._d[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

        def get(self, name):
            return self._d[name]

    def c2_msra_fill(m):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)

    import json as _json

    mod("fvcore"); mod("fvcore.common"); mod("fvcore.nn")
    mod("fvcore.common.registry", Registry=Registry)
    mod("fvcore.nn.weight_init", c2_msra_fill=c2_msra_fill)
    mod("fvcore.common.file_io", PathManager=object())
    sys.modules["simplejson"] = _json
    mod("detectron2"); mod("detectron2.layers", ROIAlign=type("ROIAlign", (nn.Module,), {}))
    mod("av")
    mod("fairseq"); mod("fairseq.search"); mod("fairseq.utils")
    dummy = lambda n: type(n, (nn.Module,), {})
    noop_deco = lambda *a, **k: (lambda f: f)
    mod("fairseq.models", FairseqIncrementalDecoder=dummy("FairseqIncrementalDecoder"),
        FairseqLanguageModel=dummy("FairseqLanguageModel"), register_model=noop_deco,
        register_model_architecture=noop_deco, ARCH_CONFIG_REGISTRY={}, ARCH_MODEL_REGISTRY={})
    mod("fairseq.models.transformer", TransformerEncoder=dummy("TransformerEncoder"),
        TransformerDecoder=dummy("TransformerDecoder"), EncoderOut=tuple,
        DEFAULT_MAX_SOURCE_POSITIONS=1024, DEFAULT_MAX_TARGET_POSITIONS=1024)
    sys.modules["fairseq"].search = sys.modules["fairseq.search"]
    sys.modules["fairseq"].utils = sys.modules["fairseq.utils"]


def load_reference():
    install_stubs()
    sys.path.insert(0, os.path.join(REF, "SlowFast"))
    sys.path.insert(0, REF)
    import vidsitu_code.mdl_sf_base as M  # noqa
    import utils.video_utils as VU  # noqa
    return M, VU


def run_case(name, M, VU):
    from common import build_model, synthetic_frames

    sf_name, n, crop, seed = CASES[name][:4]
    overrides = CASES[name][4] if len(CASES[name]) > 4 else None
    raw = "rawinit" in name
    mine, cfg, comm = build_model(sf_name, seed=seed, randomize_bn=not raw, crop=crop, sf_overrides=overrides)
    sd = mine.state_dict()
    ref = M.SFBase(cfg, comm).eval()
    missing, unexpected = ref.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert set(ref.state_dict().keys()) == set(sd.keys())
    sfc = cfg.sf_mdl
    frames = synthetic_frames(n, sfc.DATA.NUM_FRAMES, crop, seed=1234 + seed)
    # exactly the reference data path: tensor_normalize -> permute -> pack_pathway_output (dat_loader.py:478-484)
    per_clip = []
    for f in frames:
        x = VU.tensor_normalize(f, sfc.DATA.MEAN, sfc.DATA.STD).permute(3, 0, 1, 2)
        per_clip.append(VU.pack_pathway_output(sfc, x))
    npw = len(per_clip[0])
    xs = [torch.stack([c[p] for c in per_clip]).float() for p in range(npw)]
    assert n % 1 == 0
    # SFBase consumes [B, 5, ...]; use B = n/5 when divisible, else pad the event axis by viewing n as (n,1)->
    # the reference's combine_first_ax only reshapes, so feed [1, n, ...] through get_feats semantics directly.
    inp = {"frms_ev_fast_tensor": xs[-1].unsqueeze(0), "vseg_idx": torch.zeros(1, dtype=torch.long)}
    if npw == 2:
        inp["frms_ev_slow_tensor"] = xs[0].unsqueeze(0)
    t0 = time.time()
    with torch.no_grad():
        fmaps = ref.forward_encoder(inp)
        pooled = ref.head(fmaps)
        head_out = pooled.permute((0, 2, 3, 4, 1))
        logits = ref.proj_head(head_out).view(n, -1)
    dt = time.time() - t0
    pooled = pooled.flatten(1)
    top5 = torch.softmax(logits, -1).sort(dim=-1, descending=True)[1][:, :5]
    out = {
        "pooled": pooled.numpy().astype(np.float32),
        "logits": logits.numpy().astype(np.float32),
        "top5": top5.numpy().astype(np.int64),
        "slow_idx": (torch.linspace(0, sfc.DATA.NUM_FRAMES - 1, sfc.DATA.NUM_FRAMES // sfc.SLOWFAST.ALPHA).long().numpy()
                     if npw == 2 else np.zeros(0, np.int64)),
        "input_sample": np.stack([x.flatten()[:: max(1, x.numel() // 4096)][:4096].numpy() for x in xs]),
    }
    for p, f in enumerate(fmaps):
        out[f"fmap{p}_shape"] = np.array(f.shape, np.int64)
        flat = f.flatten()
        out[f"fmap{p}_sample"] = flat[:: max(1, flat.numel() // 8192)][:8192].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    meta = {"case": name, "sf_mdl_name": sf_name, "clips": n, "crop": crop, "seed": seed, "raw_init": raw,
            "sf_overrides": overrides,
            "ref_cpu_seconds": round(dt, 3), "threads": torch.get_num_threads(), "num_state_dict_keys": len(sd),
            "pooled_absmax": float(pooled.abs().max()), "torch": torch.__version__}
    print(json.dumps(meta))
    return meta


def main():
    names = sys.argv[1:] or list(CASES)
    M, VU = load_reference()
    metas = {}
    meta_path = os.path.join(HERE, "golden_meta.json")
    if os.path.exists(meta_path):
        metas = json.load(open(meta_path))
    for nm in names:
        metas[nm] = run_case(nm, M, VU)
        json.dump(metas, open(meta_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
