#!/usr/bin/env bash
# One lean GPU-box visit (a few minutes): parity tests, smoke, bench (+ per-launch table), the ncu launch
# list of one step and an `ncu --set full` pass over a SHORT list of representative launches (a whole step
# under --set full takes > 15 min and its report is > 64 MiB: never do that).
#   bash tools/gpu_round.sh [quick]      quick = tests + bench only
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
grep -B2 -A40 "^___" gpurun_out/pytest.log | head -150
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --per-op gpurun_out/per_op.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
[ "$1" = "quick" ] && exit 0
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-range > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
OPS="${VSB_FULL_OPS:-s1.pathway1_stem.conv s1.pathway0_stem.conv s2.pathway0_res1.branch2.b s2.pathway0_res1.branch2.c s2.pathway1_res1.branch2.b s3.pathway0_res1.branch2.b s4.pathway0_res1.branch2.a s4.pathway0_res1.branch2.c s5.pathway0_res1.branch2.b}"
timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/ops_full \
  python tools/profile_ops.py $OPS > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
grep profiling gpurun_out/ncu_full.log | awk '{print $2}' > gpurun_out/ops_full.labels
python tools/ncu_summary.py full gpurun_out/ops_full.ncu-rep $(cat gpurun_out/ops_full.labels) > gpurun_out/ops_full.txt 2>&1
ls -la gpurun_out
# keep the merge-back under 64 MiB
find gpurun_out -name '*.ncu-rep' -size +40M -delete
