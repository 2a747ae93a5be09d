#!/bin/bash
# one change to a factor/solve kernel: HSS + Schur suites, then the bench step twice
mkdir -p gpurun_out
T=${1:-r3d}
(timeout 400 python -m pytest tests/test_hss_gpu.py tests/test_schur_gpu.py tests/test_float_gpu.py -q -x 2>&1 | tail -n 4) | cut -c1-200
for v in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],3), 'qr_ms', round(d['roofline']['kernel_ms'],3), 'rest', round(d['ms_per_step']-d['roofline']['kernel_ms'],3), 'resid', d['config']['solve_residual'])"
done 2>&1 | tee gpurun_out/${T}_ab.log
