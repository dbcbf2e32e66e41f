python tools/sweep_variant.py base
VSB_EPI_BUFS=2 python tools/sweep_variant.py eb2
VSB_EPI_N=32 python tools/sweep_variant.py en32
VSB_EPI_N=32 VSB_EPI_BUFS=2 python tools/sweep_variant.py en32eb2
VSB_EPI_N=32 VSB_EPI_BUFS=4 python tools/sweep_variant.py en32eb4
python tools/sweep_compare.py
