#!/bin/bash
# Build libsfb200.so for sm_100a (nvcc cross-compiles without a GPU).  Output stays in-tree so it travels with gpurun.
set -e
cd "$(dirname "$0")"
SRC=shapeformer_b200/csrc
OUT=shapeformer_b200/lib
mkdir -p $OUT build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
pids=()
for f in ar_kernels ar_engine ar_chain attn_grouped enc_kernels decoder_kernels decoder_tc tc_gemm tc_gemm_ps tc_big conv_tc attn_prefill_tc mesh_kernels capi; do
  if [ ! -f build/$f.o ] || [ $SRC/$f.cu -nt build/$f.o ] || [ -n "$(find $SRC include -name '*.cuh' -newer build/$f.o -o -name '*.h' -newer build/$f.o)" ]; then
    ( $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libsfb200.so build/ar_kernels.o build/ar_engine.o build/ar_chain.o build/attn_grouped.o build/enc_kernels.o build/decoder_kernels.o build/decoder_tc.o build/tc_gemm.o build/tc_gemm_ps.o build/tc_big.o build/conv_tc.o build/attn_prefill_tc.o build/mesh_kernels.o build/capi.o -lcudart
echo "built $OUT/libsfb200.so"
