#!/bin/bash
# quick GPU check while tuning the particle kernels: engine parity tests + a short bench with the fused-stage profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_engine.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --fused-profile ${BENCH_ARGS} > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
tail -3 gpurun_out/bench_q.err
python tools/show_bench.py gpurun_out/bench_q.json | grep -v "roofline\|clocks"
python -c "import json;print(json.load(open('gpurun_out/bench_q.json'))['fused_stages'])"
