#!/usr/bin/env bash
# Round-2 closing GPU visit (after the warp-MMA thin bottleneck, the table-free pack and the new average pool): tests,
# smoke, bench (+ per-op, gpu_baseline), the other BASELINE configs, parity report, ncu launch list + DRAM bytes of one
# step, ncu --set full of the thin bottleneck kernel (res2 identity, res3 identity, res2 projection), sanitizer.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --per-op gpurun_out/per_op.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for m in i3d_r50_8x8 i3d_r50_nl_8x8; do
  timeout 600 python bench.py --model $m --steps 30 > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "bench $m rc=$?"; cut -c1-300 gpurun_out/bench_$m.json
done
timeout 600 python bench.py --model slow_fast_r101_16x8 --batch 32 --steps 20 --no-cpu-baseline > gpurun_out/bench_sf101_b32.json 2> gpurun_out/bench_sf101.err; echo "bench sf101 rc=$?"; cut -c1-300 gpurun_out/bench_sf101_b32.json
timeout 900 python tools/gpu_parity_report.py > gpurun_out/parity_report.log 2>&1; echo "parity rc=$?"; grep -c . gpurun_out/parity_report.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --profile-range > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/step_traffic.py gpurun_out/launches.csv > gpurun_out/step_dram_traffic.json; head -c 600 gpurun_out/step_dram_traffic.json
python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launches.txt; head -20 gpurun_out/launches.txt
timeout 300 python tools/gpu_check_fused.py --only thin --bench --out gpurun_out/check_thin.json > gpurun_out/check_thin.log 2>&1; echo "check thin rc=$?"; tail -1 gpurun_out/check_thin.log
for c in bench_thin_s2 bench_thin_s3 bench_thin_proj_s2_x16; do
  rm -f gpurun_out/tmp_ncu.json
  timeout 250 ncu --set full --import-source on --clock-control none -k regex:bottleneck_thin -s 3 -c 1 -f -o gpurun_out/thin_$c \
    python tools/gpu_check_fused.py --child --bench --only $c --out gpurun_out/tmp_ncu.json > gpurun_out/ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
for c in bench_thin_s2 bench_thin_s3 bench_thin_proj_s2_x16; do python tools/ncu_summary.py full gpurun_out/thin_$c.ncu-rep $c; done > gpurun_out/ncu_full_thin_bottleneck.txt; head -c 1200 gpurun_out/ncu_full_thin_bottleneck.txt
bash tools/gpu_sanitize.sh > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?"; tail -30 gpurun_out/sanitize.log
ls -la gpurun_out | tail -30
