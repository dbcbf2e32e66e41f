"""Generates tests/golden/yaml_presets.json: the hot-path keys of the reference's backbone YAMLs, read from the files
themselves (run here, where /root/reference exists; the fixture is what travels).  tests/test_host.py pins
vidsitu_b200.config.SF_MDL_PRESETS to it.

    python tests/golden/make_yaml_presets.py
"""
import json
import os

import yaml

REF = "/root/reference"
FILES = {   # mdl.sf_mdl_name -> YAML (vidsitu_code/extended_config.py:14-20, plus BASELINE.json config 4)
    "slow_fast_nl_r50_8x8": "configs/vsitu_mdl_cfgs/Kinetics_c2_SLOWFAST_8x8_R50.yaml",
    "slow_nl_r50_8x8": "configs/vsitu_mdl_cfgs/Kinetics_c2_SLOW_8x8_R50.yaml",
    "c2d_r50_8x8": "configs/vsitu_mdl_cfgs/Kinetics_C2D_8x8_R50.yaml",
    "i3d_r50_8x8": "configs/vsitu_mdl_cfgs/Kinetics_c2_I3D_8x8_R50.yaml",
    "i3d_r50_nl_8x8": "configs/vsitu_mdl_cfgs/Kinetics_c2_I3D_NLN_8x8_R50.yaml",
    "slow_fast_r101_16x8": "SlowFast/configs/Kinetics/c2/SLOWFAST_16x8_R101_50_50.yaml",
}
SECTIONS = ("DATA", "RESNET", "NONLOCAL", "MODEL", "SLOWFAST", "BN")
DROP = {("DATA", "PATH_TO_DATA_DIR"), ("DATA", "TRAIN_JITTER_SCALES"), ("DATA", "TRAIN_CROP_SIZE"),
        ("DATA", "TEST_CROP_SIZE"), ("MODEL", "LOSS_FUNC"), ("BN", "USE_PRECISE_STATS"), ("BN", "NUM_BATCHES_PRECISE"),
        ("BN", "MOMENTUM"), ("BN", "WEIGHT_DECAY")}

out = {}
for name, rel in FILES.items():
    y = yaml.safe_load(open(os.path.join(REF, rel)))
    out[name] = {"file": rel, "keys": {s: {k: v for k, v in (y.get(s) or {}).items() if (s, k) not in DROP}
                                       for s in SECTIONS if y.get(s)}}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "yaml_presets.json"), "w"), indent=1,
          sort_keys=True)
print({k: sorted(v["keys"]) for k, v in out.items()})
