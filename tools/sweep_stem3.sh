python tools/sweep_variant.py base
python tools/sweep_variant.py stem1sm '{"s1.pathway0_stem.conv": {"flags": 32}}'
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/sweep_*.json')):
    d=json.load(open(p)); o=dict(d['ops'])
    print(d['name'], 'step', round(d['step_ms'],3), 'slow stem', o['s1.pathway0_stem.conv'], 'fast stem', o['s1.pathway1_stem.conv'])
P
