#!/bin/bash
# build chimera_b200/libv_<name>.so with extra -D flags on particles_fused.cu (tuning builds, see tools/tune_fused.sh)
name=$1; shift
cd chimera_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin /usr/bin/g++ "$@" -c particles_fused.cu -o /tmp/pf_$name.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libv_$name.so particles.o particles_sorted.o /tmp/pf_$name.o spectral.o diagnostics.o gemm.o fbops.o api_host.o engine.o -lcufft -Xlinker -rpath,/usr/local/cuda/lib64
