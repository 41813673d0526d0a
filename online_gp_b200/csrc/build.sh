#!/bin/bash
# Builds libwiski_b200.so for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v"
OBJS=""
for f in interp kron kron_fused kron_tc panel gemm_tc; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ common.cuh -nt $f.o ] || [ tc_ptx.cuh -nt $f.o ] || [ ../../include/wiski_b200.h -nt $f.o ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS -c $f.cu -o $f.o 2> $f.ptxas.log || { cat $f.ptxas.log; exit 1; }
  fi
  OBJS="$OBJS $f.o"
done
$NVCC -shared -o libwiski_b200.so $OBJS -lcudart -lcuda
echo "built $(pwd)/libwiski_b200.so"
# standalone tensor-core GEMM check (run on the GPU box: online_gp_b200/csrc/test_gemm_tc)
if [ test_gemm_tc.cu -nt test_gemm_tc ] || [ gemm_tc.o -nt test_gemm_tc ]; then
  $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a test_gemm_tc.cu gemm_tc.o interp.o -o test_gemm_tc -lcudart -lcuda
fi
if [ test_kron_tc.cu -nt test_kron_tc ] || [ kron_tc.o -nt test_kron_tc ] || [ kron_fused.o -nt test_kron_tc ]; then
  $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a test_kron_tc.cu kron_tc.o kron_fused.o interp.o -o test_kron_tc -lcudart -lcuda
fi
