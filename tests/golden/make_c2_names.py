"""Generates tests/golden/c2_name_pairs.json: Caffe2 blob names -> PyTorch keys as converted by the
UNMODIFIED reference table (SlowFast/slowfast/utils/c2_model_loading.py, imported from /root/reference by
file path -- it only needs `re`).  Runs only in the build container; the fixture is committed.

The blob names are synthesised by inverting the naming scheme for every tensor of every backbone this
package builds (SlowFast-R50/R101, I3D(+NLN), Slow, C2D), plus the solver-state blobs a real Caffe2
checkpoint also carries (`*_momentum`, `lr`, `model_iter`), which must NOT land on a model key.
"""
import importlib.util
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

spec = importlib.util.spec_from_file_location(
    "c2_model_loading", "/root/reference/SlowFast/slowfast/utils/c2_model_loading.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
convert = ref.get_name_convert_func()

BN = {"weight": "s", "bias": "b", "running_mean": "rm", "running_var": "riv"}
PL = {"weight": "w", "bias": "b"}


def c2_name(key: str):
    """Inverse of the naming scheme (test data only)."""
    if key.endswith("num_batches_tracked"):
        return None
    m = re.match(r"s(\d)\.pathway(\d)_res(\d+)\.branch(1|2)(?:\.([abc]))?(_bn)?\.(\w+)$", key)
    if m:
        st, p, blk, br, abc, bn, leaf = m.groups()
        t = "t_" if p == "1" else ""
        return f"{t}res{st}_{blk}_branch{br}{abc or ''}_{'bn_' + BN[leaf] if bn else PL[leaf]}"
    m = re.match(r"s1\.pathway(\d)_stem\.(conv|bn)\.(\w+)$", key)
    if m:
        p, kind, leaf = m.groups()
        t = "t_" if p == "1" else ""
        return f"{t}res_conv1_bn_{BN[leaf]}" if kind == "bn" else f"{t}conv1_{PL[leaf]}"
    m = re.match(r"s(\d)_fuse\.(conv_f2s|bn)\.(\w+)$", key)
    if m:
        st, kind, leaf = m.groups()
        base = "t_pool1_subsample" if st == "1" else f"t_res{st}_{ {'2': 2, '3': 3, '4': 5}[st] }_branch2c_bn_subsample"
        return f"{base}_bn_{BN[leaf]}" if kind == "bn" else f"{base}_{PL[leaf]}"
    m = re.match(r"s(\d)\.pathway0_nonlocal(\d+)\.(conv_(theta|phi|g|out)|bn)\.(\w+)$", key)
    if m:
        st, idx, kind, part, leaf = m.groups()
        return f"nonlocal_conv{st}_{idx}_bn_{BN[leaf]}" if kind == "bn" else f"nonlocal_conv{st}_{idx}_{part}_{PL[leaf]}"
    m = re.match(r"head\.projection\.(\w+)$", key)
    if m:
        return f"pred_{PL[m.group(1)]}"
    raise KeyError(key)


def main():
    from common import build_model
    out = {}
    for name in ("slow_fast_nl_r50_8x8", "slow_fast_r101_16x8", "i3d_r50_8x8", "i3d_r50_nl_8x8", "slow_nl_r50_8x8",
                 "c2d_r50_8x8"):
        model, _, _ = build_model(name, seed=0, crop=64)
        keys = list(model.sf_mdl.state_dict().keys())
        pairs = []
        for k in keys:
            c2 = c2_name(k)
            if c2 is None:
                continue
            got = convert(c2)
            assert got == k, (k, c2, got)          # the reference table maps the synthesised name back
            pairs.append([c2, got])
        solver = ["lr", "model_iter", "conv1_w_momentum", "res2_0_branch2a_w_momentum", "res_conv1_bn_s_momentum",
                  "t_res4_5_branch2c_bn_subsample_w_momentum", "pred_w_momentum", "pred_b_momentum",
                  "res_conv1_bn_b_momentum", "nonlocal_conv3_1_theta_w_momentum"]
        junk = [[s, convert(s)] for s in solver]
        assert not any(j[1] in set(keys) for j in junk)
        out[name] = {"pairs": pairs, "solver_blobs": junk, "n_state_dict_keys": len(keys)}
        print(name, len(pairs), "model blobs")
    json.dump(out, open(os.path.join(HERE, "c2_name_pairs.json"), "w"))


if __name__ == "__main__":
    main()
