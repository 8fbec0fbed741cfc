#!/bin/bash
# time the fused particle kernel for each tuning build chimera_b200/libv_*.so (CHIMERA_B200_LIB override)
mkdir -p gpurun_out
for lib in chimera_b200/libv_*.so; do
  n=$(basename $lib .so)
  CHIMERA_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --ppc ${PPC:-48} > gpurun_out/tune_$n.json 2> gpurun_out/tune_$n.err
  python - "$n" gpurun_out/tune_$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    s = d["stages"]
    print("%-12s step %.2f ms  fused %.3f ms  value %.3e" % (sys.argv[1], d["ms_per_step"], s.get("particles_fused", {}).get("ms_per_call", float("nan")), d["value"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
