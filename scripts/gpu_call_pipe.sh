#!/bin/bash
# A/B of the cp.async-streamed solve sweeps (SB200_SOLVE_PIPE): parity tests with the switch on, then bench both ways
mkdir -p gpurun_out
(time SB200_SOLVE_PIPE=1 timeout 600 python -m pytest tests/test_hss_gpu.py tests/test_schur_gpu.py tests/test_dist_gpu.py -q -m gpu -x) > gpurun_out/pipe_pytest.log 2>&1
tail -6 gpurun_out/pipe_pytest.log | cut -c1-200
for v in 0 1; do
  SB200_SOLVE_PIPE=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/pipe_bench_$v.json 2> gpurun_out/pipe_bench_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/pipe_bench_$v.json").read().strip().splitlines()[-1])
print("SOLVE_PIPE=$v ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"])
PY
done
SB200_SOLVE_PIPE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/pipe_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/pipe_launches.log 2>&1
wc -l gpurun_out/pipe_launches.csv
