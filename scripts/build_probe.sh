#!/bin/bash
# Development aid: libsfb200_probe.so = the normal objects + ar_chain.cu built with -DSFB_CHAIN_PROBE=1 (timeline stamps).
# Use with SFB200_LIB=shapeformer_b200/lib/libsfb200_probe.so python scripts/chain_timeline.py
set -e
cd "$(dirname "$0")/.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DSFB_CHAIN_PROBE=1 -c shapeformer_b200/csrc/ar_chain.cu -o build/ar_chain_probe.o
objs=""
for f in build/*.o; do case $f in build/ar_chain.o|build/ar_chain_probe.o) ;; *) objs="$objs $f";; esac; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o shapeformer_b200/lib/libsfb200_probe.so $objs build/ar_chain_probe.o -lcudart
echo built shapeformer_b200/lib/libsfb200_probe.so
