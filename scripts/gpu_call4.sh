#!/bin/bash
mkdir -p gpurun_out
T=${1:-c4}
timeout 600 python scripts/variant_check.py SB200_QR_NOWIDE=1 SB200_QR_VARIANT=1 SB200_QR_VARIANT=2 > gpurun_out/${T}_check.log 2>&1; tail -18 gpurun_out/${T}_check.log
for cfg in "SB200_QR_NOWIDE=1" "SB200_QR_VARIANT=0" "SB200_QR_VARIANT=1" "SB200_QR_VARIANT=2"; do echo "$cfg bench"; env $cfg timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_$cfg.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
for v in 0 1; do SB200_QR_VARIANT=$v timeout 300 python scripts/qr_timing.py 262144 > gpurun_out/${T}_timing_v$v.log 2>&1; grep "block 0 .*m 256" gpurun_out/${T}_timing_v$v.log | sort | head -8; done
