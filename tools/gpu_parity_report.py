"""Prints parity metrics of every golden case (bf16 and fp32 modes) -- run on the GPU box."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import build_model, synthetic_frames
GOLD = os.path.join(ROOT, "tests", "golden")
META = json.load(open(os.path.join(GOLD, "golden_meta.json")))
from oracle import sf_oracle as O   # checker (this is a test tool)


def metrics(f, r, lg, g):
    err = np.abs(f - r)
    cos = float(((f * r).sum(-1) / (np.linalg.norm(f, axis=-1) * np.linalg.norm(r, axis=-1))).min())
    mean = float(np.abs(r).mean())
    top5 = np.sort(np.argsort(-lg, axis=-1, kind="stable")[:, :5], -1)
    return {"cos": cos, "max_abs_over_max_ref": float(err.max() / np.abs(r).max()),
            "rel_floor_0.1mean": float((err / np.maximum(np.abs(r), 0.1 * mean)).max()),
            "rel_floor_mean": float((err / np.maximum(np.abs(r), mean)).max()),
            "rel_l2": float(np.linalg.norm(f - r) / np.linalg.norm(r)),
            "p99_rel_floor_0.1mean": float(np.percentile(err / np.maximum(np.abs(r), 0.1 * mean), 99)),
            "logit_cos": float(((lg * g["logits"]).sum(-1) / (np.linalg.norm(lg, axis=-1) * np.linalg.norm(g["logits"], axis=-1))).min()),
            "top5_set_equal": bool(np.array_equal(top5, np.sort(g["top5"], -1)))}


from common import torch_bf16_comparator  # noqa: E402


out = {}
cases = sys.argv[1:] or list(META)
for case in cases:
    m = META[case]; g = np.load(os.path.join(GOLD, case + ".npz"))
    for prec in ("bf16", "fp32"):
        model, cfg, _ = build_model(m["sf_mdl_name"], seed=m["seed"], randomize_bn=not m["raw_init"], crop=m["crop"], precision=prec,
                                    sf_overrides=m.get("sf_overrides"))
        model = model.cuda()
        frames = synthetic_frames(m["clips"], cfg.sf_mdl.DATA.NUM_FRAMES, m["crop"], seed=1234 + m["seed"]).cuda()
        feats, logits = model.extract_features(frames, want_logits=True)
        torch.cuda.synchronize()
        f, r = feats.cpu().numpy(), g["pooled"]
        err = np.abs(f - r)
        cos = float(((f * r).sum(-1) / (np.linalg.norm(f, axis=-1) * np.linalg.norm(r, axis=-1))).min())
        mean = float(np.abs(r).mean())
        lg = logits.cpu().numpy()
        top5 = np.sort(np.argsort(-lg, axis=-1, kind="stable")[:, :5], -1)
        rec = {"cos": cos, "max_abs_over_max_ref": float(err.max() / np.abs(r).max()),
               "rel_floor_0.1mean": float((err / np.maximum(np.abs(r), 0.1 * mean)).max()),
               "rel_floor_mean": float((err / np.maximum(np.abs(r), mean)).max()),
               "rel_l2": float(np.linalg.norm(f - r) / np.linalg.norm(r)),
               "p99_rel_floor_0.1mean": float(np.percentile(err / np.maximum(np.abs(r), 0.1 * mean), 99)),
               "mean_ref": mean, "max_ref": float(np.abs(r).max()),
               "logit_cos": float(((lg * g["logits"]).sum(-1) / (np.linalg.norm(lg, axis=-1) * np.linalg.norm(g["logits"], axis=-1))).min()),
               "top5_set_equal": bool(np.array_equal(top5, np.sort(g["top5"], -1))),
               "logit_gap_5_6": float(np.min(-np.sort(-g["logits"], -1)[:, 4] + np.sort(-g["logits"], -1)[:, 5] * -1)) }
        out[f"{case}/{prec}"] = rec
        print(case, prec, json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
        if prec == "bf16":
            try:
                cf, cl = torch_bf16_comparator(model, cfg, frames.cpu())
                crec = metrics(cf, g["pooled"], cl, g)
                out[f"{case}/torch_bf16_channels_last_3d"] = crec
                print(case, "torch_bf16_cl3d", json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in crec.items()}), flush=True)
            except Exception as e:  # noqa: BLE001
                out[f"{case}/torch_bf16_channels_last_3d"] = {"error": str(e)[:200]}
                print(case, "torch_bf16_cl3d error", str(e)[:200], flush=True)
        del model
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
