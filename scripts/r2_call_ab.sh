#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2j}
(time timeout 600 python -m pytest tests/test_hss_gpu.py tests/test_compress_gpu.py tests/test_schur_gpu.py -q -m gpu -x 2>&1 | tail -n 25) > gpurun_out/${T}_pytest.log 2>&1
tail -n 6 gpurun_out/${T}_pytest.log | cut -c1-220
for v in 1 0; do
  SB200_FWD_RL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_fwdrl$v.json 2> gpurun_out/${T}_bench_fwdrl$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_fwdrl$v.json"))
print("FWD_RL=$v ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"])
PY
done
