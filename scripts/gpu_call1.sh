#!/bin/bash
# round-1 re-entry verification: GPU tests, LL A/B, phase timing, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/c1_pytest.log 2>&1
tail -3 gpurun_out/c1_pytest.log
timeout 600 python scripts/ll_check.py > gpurun_out/c1_ll_check.log 2>&1; tail -20 gpurun_out/c1_ll_check.log
for v in 0 1; do echo "LL=$v bench"; SB200_QR_LL=$v timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/c1_bench_ll$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
for v in 0 1; do SB200_QR_LL=$v timeout 300 python scripts/qr_timing.py 262144 > gpurun_out/c1_timing_ll$v.log 2>&1; grep -m 10 timing gpurun_out/c1_timing_ll$v.log; done
