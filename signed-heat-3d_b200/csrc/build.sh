#!/bin/bash
# Build libshm3d_grid.so (sm_100a only).  Usage: csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NCCL_INC=${NCCL_INC:-/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/include}
OUT=../lib/libshm3d_grid.so
mkdir -p ../lib
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC,-pthread,-mavx2,-mfma,-Wall \
     -I"$NCCL_INC" -shared -o $OUT \
     k_sum.cu grid_ops.cu projector.cu sources.cu dist.cu solver.cu host_api.cu \
     -ldl -lpthread "$@"
echo "built $OUT"
