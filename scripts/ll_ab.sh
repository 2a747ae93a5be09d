#!/bin/bash
# A/B of the left-looking leaf QR: parity of the factors, the GPU tests, and the bench
timeout 600 python scripts/ll_check.py 2>&1 | tail -20
for v in 1 2; do echo "LL=$v tests"; SB200_QR_LL=$v timeout 900 python -m pytest tests/test_hss_gpu.py tests/test_compress_gpu.py -x -q -m gpu 2>&1 | tail -2; done
for v in 0 1 2; do echo "LL=$v bench"; SB200_QR_LL=$v timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
