#!/usr/bin/env bash
# Builds vidsitu_b200/libvidsitu_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libvidsitu_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "${here}/build"
objs=()
for src in api conv_igemm_sm100 conv_igemm2_sm100 conv_win_sm100 bottleneck_fused_sm100 bottleneck_thin_sm100 conv_simt ops_mem topk nonlocal probe_umma program stem_pool_sm100 jpeg_ingest; do
  obj="${here}/build/${src}.o"
  if [[ ! -f "${obj}" || "${here}/${src}.cu" -nt "${obj}" || "${here}/ptx.cuh" -nt "${obj}" || "${here}/common.h" -nt "${obj}" || "${here}/conv_plan.h" -nt "${obj}" || "${here}/epilogue.cuh" -nt "${obj}" || "${here}/bottleneck_thin.h" -nt "${obj}" || "${here}/../../include/vidsitu_b200.h" -nt "${obj}" ]]; then
    "${NVCC}" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC -Xptxas -v -c "${here}/${src}.cu" -o "${obj}" 2> "${here}/build/${src}.ptxas.log" \
      || { cat "${here}/build/${src}.ptxas.log" >&2; exit 1; }
  fi
  objs+=("${obj}")
done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "${out}" "${objs[@]}"
echo "built ${out}"
