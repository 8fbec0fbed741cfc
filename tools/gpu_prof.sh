#!/bin/bash
# ncu evidence for the bench command: launch list (share of step) + --set full captures of the hot kernels.
# gpurun brings back at most 64 MiB: the full captures are kept to ~10 launches each (~2.7 MB per launch).
mkdir -p gpurun_out
ARGS="--steps ${STEPS:-12} --warmup ${WARMUP:-0} --no-e2e --no-cpu --ppc ${PPC:-48}"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launch_bench.log 2>&1
# the step kernel and the contraction launches of one fused step (the first step of a call and its 15 contractions are skipped)
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:fused_particles_k|gemm_dmma_k" -s ${SKIP:-16} -c ${COUNT:-9} -f -o gpurun_out/prof python bench.py $ARGS > gpurun_out/ncu_full_bench.log 2>&1
# the two halves of the re-binning step (step 11): gather + push + push_coords before the sort, dep_curr + dep_dens after it
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:gather_push_coords_k|fused_pass_k" -s 1 -c 2 -f -o gpurun_out/prof_rebin python bench.py $ARGS > gpurun_out/ncu_full2_bench.log 2>&1
ls -la gpurun_out/ | tail -6
