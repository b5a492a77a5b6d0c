#!/usr/bin/env bash
# Builds libmetro.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libmetro.so
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
pids=()
for f in conv_gemm conv_chain softargmax root_fused post crops strict metro_api; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o &
  pids+=($!)
done
g++ -O2 -std=c++17 -fPIC -I/usr/local/cuda/include -c plan.cpp -o build/plan.o &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/conv_gemm.o build/conv_chain.o build/softargmax.o build/root_fused.o build/post.o build/crops.o build/strict.o build/metro_api.o build/plan.o -cudart static
echo "built $(readlink -f $OUT)"
