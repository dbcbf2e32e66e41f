python tools/sweep_variant.py base
python tools/sweep_variant.py stem2sm '{"s1.pathway0_stem.conv": {"flags": 16}}'
VSB_PROFILE_MODEL=i3d_r50_8x8 python tools/sweep_variant.py i3d_base
VSB_PROFILE_MODEL=i3d_r50_8x8 python tools/sweep_variant.py i3d_stem2sm '{"s1.pathway0_stem.conv": {"flags": 16}}'
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/sweep_*.json')):
    d=json.load(open(p)); o=dict(d['ops'])
    print(d['name'], 'step', round(d['step_ms'],3), 'stem', o['s1.pathway0_stem.conv'])
P
