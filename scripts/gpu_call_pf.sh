#!/bin/bash
# A/B: streamed (cp.async) solve sweeps vs the non-streamed kernels with L2 prefetch of the next block
mkdir -p gpurun_out
(SB200_SOLVE_PIPE=12 timeout 300 python -m pytest tests/test_hss_gpu.py tests/test_schur_gpu.py -q -m gpu -x) > gpurun_out/pf_pytest.log 2>&1
tail -n 2 gpurun_out/pf_pytest.log | cut -c1-200
for v in 3 6 9 12; do
  SB200_SOLVE_PIPE=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/pf_bench_$v.json 2> gpurun_out/pf_bench_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/pf_bench_$v.json").read().strip().splitlines()[-1])
print("SOLVE_PIPE=$v ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"])
PY
done
