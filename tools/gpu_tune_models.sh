set -x
mkdir -p gpurun_out
export VSB_TUNE_ONLY=en32,en32eb3,en32eb4,st2,st3,bn128,bn256,stream_w,st2en32,st3en32,st2bn128,bn256en32eb4,bn256en32,st2sw,en32eb4sw,sm1,sm2,sm2en32,en16
python tools/autotune.py --model i3d_r50_nl_8x8 --clips 64 --write 2>&1 | tail -3
python tools/autotune.py --model slow_fast_r101_16x8 --clips 32 --write 2>&1 | tail -3
cp vidsitu_b200/tune_table.json gpurun_out/
VSB_PROFILE_MODEL=i3d_r50_nl_8x8 python tools/sweep_variant.py i3dnl | tail -1
VSB_PROFILE_MODEL=slow_fast_r101_16x8 VSB_PROFILE_CLIPS=32 python tools/sweep_variant.py sf101 | tail -1
