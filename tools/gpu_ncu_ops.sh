# ncu --set full of a short list of launches (names: tools/profile_ops.py) -> gpurun_out/$1.txt (+ .ncu-rep)
#   bash tools/gpu_ncu_ops.sh NAME op [op ...]
name=$1; shift
mkdir -p gpurun_out
timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/$name \
  python tools/profile_ops.py "$@" > gpurun_out/$name.log 2>&1; echo "ncu rc=$?"
grep "^profiling" gpurun_out/$name.log | awk '{print $2}' > gpurun_out/$name.labels
python tools/ncu_summary.py full gpurun_out/$name.ncu-rep $(cat gpurun_out/$name.labels) > gpurun_out/$name.txt 2>&1
grep -E "^## |duration|tensor pipe|DRAM throughput|L2 throughput" gpurun_out/$name.txt
find gpurun_out -name '*.ncu-rep' -size +40M -delete
