set -x
python tools/sweep_variant.py base
python tools/sweep_variant.py bn128 '{"*": {"block_n": 128}}'
python tools/sweep_variant.py bn64 '{"*": {"block_n": 64}}'
VSB_EPI_WARPS=16 python tools/sweep_variant.py ew16
VSB_NO_BRES=1 python tools/sweep_variant.py nobres
VSB_WIN_ONE_CTA=1 python tools/sweep_variant.py win1cta
VSB_WIN_NO_PAIR=1 python tools/sweep_variant.py nopair
python tools/sweep_variant.py st4 '{"*": {"stages": 4}}'
python tools/sweep_variant.py st6 '{"*": {"stages": 6}}'
python tools/sweep_variant.py im2col '{"*": {"algo": "im2col"}}'
python tools/sweep_variant.py 1stream '{"*": {"streams": 1}}'
