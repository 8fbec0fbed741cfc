#!/bin/bash
# one GPU session: parity tests, smoke, bench; logs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
for PPC in ${PPCS:-16 48}; do
  timeout 900 python bench.py --ppc $PPC --steps ${STEPS:-10} --warmup 3 $BENCH_ARGS > gpurun_out/bench_ppc$PPC.json 2> gpurun_out/bench_ppc$PPC.err
  tail -5 gpurun_out/bench_ppc$PPC.err; cat gpurun_out/bench_ppc$PPC.json
done
