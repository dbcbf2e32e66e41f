#!/usr/bin/env bash
# A/B of an environment knob on the headline bench (value only), interleaved so clock drift cancels:
#   tools/gpu_ab.sh VSB_PDL 2 0 1          (knob, repeats, values...)
knob="$1"; reps="$2"; shift 2
mkdir -p gpurun_out
for i in $(seq 1 "$reps"); do
  for v in "$@"; do
    tag=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
    env "$knob=$v" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-gpu-baseline \
      2> gpurun_out/ab_${knob}_${tag}_$i.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$knob=$v run $i: value', d['value'], 'ms', d['ms_per_step'], 'conv_ms', d['roofline'].get('conv_ms_per_step'), 'clk', d['clocks']['sm_mhz'], 'launches', d.get('launches_per_step'))
" | tee -a gpurun_out/ab_${knob}.log
  done
done
