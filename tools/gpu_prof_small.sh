#!/bin/bash
# ncu evidence for a demo-size config (bench.py --config): launch list + --set full of a few contraction launches
mkdir -p gpurun_out
CFG=${CFG:-c2_pic}
ARGS="--config $CFG --steps 4 --warmup 3 --no-e2e --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$CFG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$CFG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_dmma_k" -s 60 -c 4 -f -o gpurun_out/prof_$CFG python bench.py $ARGS > gpurun_out/ncu_full_$CFG.log 2>&1
ls -la gpurun_out/ | tail -6
