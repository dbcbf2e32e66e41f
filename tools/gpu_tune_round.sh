set -x
mkdir -p gpurun_out
VSB_TUNE_ONLY=en32,en32eb3,en32eb4,st2,st3,bn128,bn256,stream_w,st2en32,st3en32,st2bn128,st3bn128,bn256en32eb4,bn256en32,st2sw,en32eb4sw,sm1,sm2,sm2en32,sm2st4 python tools/autotune.py --write 2>&1 | tail -4
cp vidsitu_b200/tune_table.json gpurun_out/
bash tools/gpu_round.sh quick
