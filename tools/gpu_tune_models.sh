# Re-measures vidsitu_b200/tune_table.json for the three benchmarked backbones (about 2.5 GPU-minutes).
set -x
mkdir -p gpurun_out
rm -f vidsitu_b200/tune_table.json
export VSB_TUNE_ONLY=en32,en32eb3,en32eb4,st2,st3,bn128,bn256,stream_w,st2en32,st3en32,st2bn128,bn256en32eb4,bn256en32,st2sw,en32eb4sw,sm1,sm2,sm2en32,en16,nots
python tools/autotune.py --model slow_fast_nl_r50_8x8 --clips 64 --write 2>&1 | tail -2
python tools/autotune.py --model i3d_r50_nl_8x8 --clips 64 --write 2>&1 | tail -2
python tools/autotune.py --model slow_fast_r101_16x8 --clips 32 --write 2>&1 | tail -2
cp vidsitu_b200/tune_table.json gpurun_out/
