set -x
python tools/sweep_variant.py base '{"*": {"early_lateral": 0}}'
python tools/sweep_variant.py el
python tools/sweep_variant.py el_p1 '{"*": {"prio": 1}}'
python tools/sweep_variant.py el_p2 '{"*": {"prio": 2}}'
