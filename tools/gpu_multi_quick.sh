#!/usr/bin/env bash
# Short multi-GPU refresh: SF101 shard check (bit-for-bit vs 1 GPU) and a bench.py line on N GPUs.
#   bash tools/gpu_multi_quick.sh 2
set -x
mkdir -p gpurun_out
N=$1
P=$((29500 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
timeout 600 $TR tools/gpu_shard_check.py --videos 16 > gpurun_out/shard_check_sf101_n$N.json 2> gpurun_out/shard_n$N.err; echo "shard N=$N rc=$?"; tail -1 gpurun_out/shard_check_sf101_n$N.json
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cut -c1-200 gpurun_out/bench_n$N.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n$N.json
