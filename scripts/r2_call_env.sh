#!/bin/bash
# bench step under an environment switch: $2 = VAR, remaining args = values
mkdir -p gpurun_out
T=${1:-r3j}; SW=$2; shift 2
for v in "$@" "$@"; do
  env $SW=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$SW=$v', 'ms', round(d['ms_per_step'],3), 'qr_ms', round(d['roofline']['kernel_ms'],3), 'rest', round(d['ms_per_step']-d['roofline']['kernel_ms'],3), 'resid', d['config']['solve_residual'])"
done 2>&1 | tee gpurun_out/${T}_ab.log
