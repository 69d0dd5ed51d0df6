#!/bin/bash
# Builds libe2b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$HERE"
SRCS="e2_api.cu e2_pool.cu e2_conv_ffma.cu e2_conv_c1.cu e2_conv_c1_tc.cu e2_conv_c1_ws.cu e2_conv_pw.cu e2_conv_tc.cu e2_conv_zstack_tc.cu e2_wgrad_tc.cu e2_wgrad_halo_tc.cu e2_wgrad_zs_tc.cu e2_wgrad_c1_tc.cu e2_train.cu e2_epilogue.cu"
OBJS=""
pids=""
for s in $SRCS; do
  o="$OUT/${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$HERE/$s" -nt "$o" ] || [ -n "$(find "$HERE" "$ROOT/include" -name '*.cuh' -newer "$o" -o -name '*.h' -newer "$o")" ]; then
    $NVCC $FLAGS ${E2_PTXAS_V:+-Xptxas -v} -c "$HERE/$s" -o "$o" &
    pids="$pids $!"
  fi
  OBJS="$OBJS $o"
done
for p in $pids; do wait $p; done
$NVCC -shared -o "$OUT/libe2b200.so" $OBJS -lcuda
echo "built $OUT/libe2b200.so"
