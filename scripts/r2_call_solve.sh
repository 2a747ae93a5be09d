#!/bin/bash
# solve-sweep A/B: HSS + Schur suites with the new default, then the bench step with the switch given in $2 at 0 / 1
mkdir -p gpurun_out
T=${1:-r3b}; SW=${2:-SB200_FWD_RING}
(timeout 400 python -m pytest tests/test_hss_gpu.py tests/test_schur_gpu.py tests/test_float_gpu.py -q -x 2>&1 | tail -n 4) | cut -c1-200
for v in 0 1 0 1; do
  env $SW=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$SW=$v', 'ms', round(d['ms_per_step'],3), 'qr_ms', round(d['roofline']['kernel_ms'],3), 'rest', round(d['ms_per_step']-d['roofline']['kernel_ms'],3), 'resid', d['config']['solve_residual'])"
done 2>&1 | tee gpurun_out/${T}_ab.log
