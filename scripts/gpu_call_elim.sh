#!/bin/bash
# A/B of the ulv_eliminate_kernel variants (SB200_ELIM_VARIANT)
mkdir -p gpurun_out
for v in 1 2; do
(SB200_ELIM_VARIANT=$v timeout 600 python -m pytest tests/test_hss_gpu.py tests/test_schur_gpu.py tests/test_configs_gpu.py -q -m gpu -x) > gpurun_out/elim_pytest_$v.log 2>&1
tail -n 2 gpurun_out/elim_pytest_$v.log | cut -c1-200
done
for v in 0 1 2; do
  SB200_ELIM_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/elim_bench_$v.json 2> gpurun_out/elim_bench_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/elim_bench_$v.json").read().strip().splitlines()[-1])
print("ELIM_VARIANT=$v ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"])
PY
done
