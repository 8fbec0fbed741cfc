#!/bin/bash
# one GPU session: parity tests, smoke, bench; logs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --ppc 16 --steps 5 --warmup 3 > gpurun_out/bench_ppc16.json 2> gpurun_out/bench_ppc16.err
tail -5 gpurun_out/bench_ppc16.err; cat gpurun_out/bench_ppc16.json
