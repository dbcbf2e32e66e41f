python tools/sweep_variant.py base
python tools/sweep_variant.py fwg2 '{"s1.pathway1_stem.conv": {"win_group": 2}}'
python tools/sweep_variant.py fwg8 '{"s1.pathway1_stem.conv": {"win_group": 8}}'
python tools/sweep_variant.py swin2 '{"s1.pathway0_stem.conv": {"algo": "window", "win_group": 2}}'
python tools/sweep_variant.py swin4 '{"s1.pathway0_stem.conv": {"algo": "window", "win_group": 4}}'
python tools/sweep_variant.py swin8 '{"s1.pathway0_stem.conv": {"algo": "window", "win_group": 8}}'
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/sweep_*.json')):
    d=json.load(open(p)); o=dict(d['ops'])
    print(d['name'], 'step', round(d['step_ms'],3), 'slow stem', o['s1.pathway0_stem.conv'], 'fast stem', o['s1.pathway1_stem.conv'])
P
