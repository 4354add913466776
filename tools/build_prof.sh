#!/bin/sh
# Development build with phase clocks / ablation switches (see CALB2_PROFILE in calfit_kernels.cuh).
# Use with CALB2_LIB=$PWD/calamity_b200/_lib/libcalamity_b200_prof.so CALB2_DBG=<bits> (128 = phase clocks).
cd "$(dirname "$0")/../calamity_b200/csrc" && nvcc -O3 -std=c++17 -DCALB2_PROFILE -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -shared -o ../_lib/libcalamity_b200_prof.so calfit_api.cu -ldl
