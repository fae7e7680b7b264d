#!/bin/bash
# Build libshm3d_grid.so (sm_100a only).  Usage: csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
# nccl.h: $NCCL_INC, else the nvidia-nccl wheel of the python on PATH, else the system include path
if [ -z "$NCCL_INC" ]; then
    NCCL_INC=$(python -c "import importlib.util, os; s = importlib.util.find_spec('nvidia.nccl'); print(os.path.join(list(s.submodule_search_locations)[0], 'include') if s else '')" 2>/dev/null || true)
fi
NCCL_INC=${NCCL_INC:-/usr/include}
OUT=../lib/libshm3d_grid.so
mkdir -p ../lib
# point_weights.cpp (row N1, host only) is compiled without floating-point contraction: its degenerate-case decisions
# (cocircular / collinear / equidistant points) must follow geometry-central's expressions bit for bit, whatever FMA the
# host compiler would like to form
mkdir -p ../build
g++ -std=c++17 -O3 -fPIC -pthread -ffp-contract=off -Wall -c point_weights.cpp -o ../build/point_weights.o
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC,-pthread,-mavx2,-mfma,-Wall \
     -I"$NCCL_INC" -shared -o $OUT \
     k_sum.cu grid_ops.cu projector.cu sources.cu dist.cu isosurface.cu solver.cu host_api.cu host_blas.cpp ../build/point_weights.o \
     -ldl -lpthread "$@"
echo "built $OUT"
# headless driver on top of the C++ mirror (include/shm3d/signed_heat_grid_solver.hpp)
mkdir -p ../bin
g++ -std=c++11 -O2 -Wall -o ../bin/shm3d_cli ../../tools/shm3d_cli.cpp -L../lib -lshm3d_grid -Wl,-rpath,'$ORIGIN/../lib'
echo "built ../bin/shm3d_cli"
