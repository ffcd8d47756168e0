#!/bin/bash
# Builds libngpb200.so in-tree for sm_100a. -fmad=false: float ops round as written (see csrc/nerf_device.cuh).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p build
pids=()
for f in common hash_grid nerf_mlp nerf_mlp_pipe umma_probe camera_optimizer model nerf_sampling nerf_loss losses optimizer density_grid render testbed; do
	if [ -f csrc/$f.cu ]; then
		( $NVCC $FLAGS -c csrc/$f.cu -o build/$f.o ) &
		pids+=($!)
	fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o libngpb200.so build/*.o -lcudart -ldl
echo "built $(pwd)/libngpb200.so"
