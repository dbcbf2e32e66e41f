#!/usr/bin/env bash
# Multi-GPU artifacts of BASELINE.json configs 4 / 5 on the GPUs of this box: for every N given on the command line,
# the SF101 shard check (bit-for-bit vs 1 GPU), the SF50 batch sweep and a short bench.py run.
#   bash tools/gpu_multi.sh 2 4
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -14 gpurun_out/topo.txt
for N in "$@"; do
  P=$((29500 + N))
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
  timeout 600 $TR tools/gpu_shard_check.py --videos 16 > gpurun_out/shard_check_sf101_n$N.json 2> gpurun_out/shard_n$N.err; echo "shard N=$N rc=$?"; tail -1 gpurun_out/shard_check_sf101_n$N.json
  timeout 600 $TR tools/gpu_sweep.py > gpurun_out/sweep_w$N.log 2>&1; echo "sweep N=$N rc=$?"; grep clips gpurun_out/sweep_w$N.log | tail -4
  timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cut -c1-200 gpurun_out/bench_n$N.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n$N.json
  if [ "$N" = "8" ]; then
    VSB_NUMA_BIND=0 timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_nobind.json 2> gpurun_out/bench_n${N}_nobind.err; echo "bench nobind rc=$?"; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n${N}_nobind.json
  fi
done
