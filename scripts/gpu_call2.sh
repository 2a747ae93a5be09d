#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/variant_check.py SB200_QR_VARIANT=1 SB200_QR_VARIANT=2 > gpurun_out/c2_check.log 2>&1; tail -20 gpurun_out/c2_check.log
for v in 0 1 2; do echo "VARIANT=$v bench"; SB200_QR_VARIANT=$v timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/c2_bench_v$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
for v in 0 1; do SB200_QR_VARIANT=$v timeout 300 python scripts/qr_timing.py 262144 > gpurun_out/c2_timing_v$v.log 2>&1; grep "m 256" gpurun_out/c2_timing_v$v.log | sort | head -8; done
SB200_QR_VARIANT=1 timeout 600 python -m pytest tests/test_hss_gpu.py tests/test_compress_gpu.py tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -2
