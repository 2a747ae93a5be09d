#!/bin/bash
mkdir -p gpurun_out
T=${1:-c6}
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
for cfg in "SB200_QR_VARIANT=1" "SB200_QR_VARIANT=3"; do echo "$cfg bench"; env $cfg timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_$cfg.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
