#!/usr/bin/env bash
# Build libdnpy_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libdnpy_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(${DNB_EXTRA_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
       -Xcompiler -fPIC)
SRCS=(elementwise.cu dense.cu dense_head.cu conv.cu conv_thin.cu pool.cu norm.cu loss.cu embed_rnn.cu rnn_tc.cu tc_gemm.cu tc_conv.cu tc_conv_sv.cu tc_conv_svp.cu tc_conv_c1p.cu tc_conv_wg.cu tc_conv_thin.cu fused_cpa.cu dp_peer.cu)
mkdir -p "$HERE/build"
pids=()
for s in "${SRCS[@]}"; do
  o="$HERE/build/${s%.cu}.o"
  if [[ ! -f "$o" || "$HERE/$s" -nt "$o" || "$HERE/common.cuh" -nt "$o" || "$HERE/tc_gemm.cuh" -nt "$o" || "$HERE/tc_conv.cuh" -nt "$o" || "$HERE/tc_common.cuh" -nt "$o" || "$HERE/../../include/dnpy_b200.h" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" ${DNB_PTXAS_V:+-Xptxas -v} -c "$HERE/$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
objs=()
for s in "${SRCS[@]}"; do objs+=("$HERE/build/${s%.cu}.o"); done
"$NVCC" -shared -o "$OUT" "${objs[@]}" -lcudart_static -Xcompiler -fPIC
echo "built $OUT"
