#!/bin/bash
# quick GPU check used during development: parity tests + one bench line summary
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --n ${1:-1048576} --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_last.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'frac',round(d['roofline']['frac'],3), 'resid',d['config']['solve_residual'], 'compress_s',round(d['config']['compress_s'],2), 'e2e_ms',round(d['e2e']['ms_per_step'],2))"
