export VSB_PROFILE_MODEL=i3d_r50_8x8
python tools/sweep_variant.py base
python tools/sweep_variant.py sm2 '{"s1.pathway0_stem.conv": {"flags": 16}}'
python tools/sweep_variant.py g2 '{"s1.pathway0_stem.conv": {"group": 2}}'
python tools/sweep_variant.py g2sm2 '{"s1.pathway0_stem.conv": {"group": 2, "flags": 16}}'
python tools/sweep_variant.py g8 '{"s1.pathway0_stem.conv": {"group": 8}}'
python tools/sweep_variant.py g8sm2 '{"s1.pathway0_stem.conv": {"group": 8, "flags": 16}}'
python tools/sweep_variant.py win2 '{"s1.pathway0_stem.conv": {"algo": "window", "win_group": 2}}'
python tools/sweep_variant.py win1 '{"s1.pathway0_stem.conv": {"algo": "window", "win_group": 1}}'
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/sweep_*.json')):
    d=json.load(open(p)); o=dict(d['ops'])
    if 's1.pathway0_stem.conv' in o and 's1.pathway1_stem.conv' not in o: print(d['name'], 'step', round(d['step_ms'],3), 'stem', o['s1.pathway0_stem.conv'])
P
