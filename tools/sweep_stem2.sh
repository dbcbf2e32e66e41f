python tools/sweep_variant.py base
python tools/sweep_variant.py sg2 '{"s1.pathway0_stem.conv": {"group": 2}}'
python tools/sweep_variant.py sg8 '{"s1.pathway0_stem.conv": {"group": 8}}'
python tools/sweep_variant.py sg4bn128 '{"s1.pathway0_stem.conv": {"block_n": 128}}'
python tools/sweep_variant.py sg4st2 '{"s1.pathway0_stem.conv": {"stages": 2}}'
python tools/sweep_variant.py sg4en32 '{"s1.pathway0_stem.conv": {"epi_n": 32, "epi_bufs": 2}}'
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/sweep_*.json')):
    d=json.load(open(p)); o=dict(d['ops'])
    print(d['name'], 'step', round(d['step_ms'],3), 'slow stem', o['s1.pathway0_stem.conv'], 'fast stem', o['s1.pathway1_stem.conv'])
P
