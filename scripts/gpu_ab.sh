#!/bin/bash
# A/B of engine builds: bench each library given on the command line (name=path)
mkdir -p gpurun_out
for kv in "$@"; do name=${kv%%=*}; lib=${kv#*=}; echo "== $name"; SB200_LIB=$PWD/$lib timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/ab_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
