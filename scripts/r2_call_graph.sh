#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2l}
(time timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15) > gpurun_out/${T}_pytest.log 2>&1
tail -n 6 gpurun_out/${T}_pytest.log | cut -c1-220
for v in 1 0; do
  SB200_GRAPH=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_graph$v.json 2> gpurun_out/${T}_bench_graph$v.err
  tail -n 3 gpurun_out/${T}_bench_graph$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_graph$v.json"))
print("GRAPH=$v ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"], "launches", d["gpu_launches"], "qr_ms", d["roofline"]["kernel_ms"], "compress_err", d["config"]["compress_rel_err"])
PY
done
