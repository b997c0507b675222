#!/bin/bash
# builds sci_b200/lib/libflomat_cuda_prof.so with -DFM_LU_PROF (clock64 phase counters in the LU panel kernel)
cd "$(dirname "$0")/../sci_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --cudart shared -DFM_LU_PROF -o ../lib/libflomat_cuda_prof.so api.cu ewise.cu norm.cu dgemm.cu lu.cu expm.cu level12.cu xfer.cu mg.cu
