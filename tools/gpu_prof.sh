#!/bin/bash
# ncu evidence for the bench command: launch list (share of step) + one --set full capture of the hot kernels
mkdir -p gpurun_out
ARGS="--steps ${STEPS:-3} --warmup ${WARMUP:-1} --no-e2e --no-cpu --ppc ${PPC:-48}"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${KERNELS:-fused_particles_k|gemm_dmma_k|gather_push_coords_k|fused_pass_k}" -s ${SKIP:-20} -c ${COUNT:-22} -f -o gpurun_out/prof python bench.py $ARGS > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out/ | tail -5
