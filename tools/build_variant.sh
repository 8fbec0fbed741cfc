#!/bin/bash
# build chimera_b200/libv_<name>.so with extra -D flags on particles_fused.cu and engine.cu (tuning builds, see
# tools/tune_fused.sh); the other objects are the regular build's
name=$1; shift
cd chimera_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin /usr/bin/g++"
nvcc $FLAGS "$@" -Xptxas -v -c particles_fused.cu -o /tmp/pf_$name.o 2> /tmp/pf_$name.log || { cat /tmp/pf_$name.log; exit 1; }
grep -A2 "fused_particles_kILi0ELi3ELi1" /tmp/pf_$name.log | grep -i "registers\|spill" | head -2
nvcc $FLAGS "$@" -c engine.cu -o /tmp/en_$name.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libv_$name.so particles.o particles_sorted.o /tmp/pf_$name.o spectral.o diagnostics.o sr.o gemm.o fbops.o staging.o api_host.o /tmp/en_$name.o -lcufft -lpthread -Xlinker -rpath,/usr/local/cuda/lib64
