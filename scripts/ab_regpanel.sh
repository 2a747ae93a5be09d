#!/bin/bash
python -m pytest tests/test_hss_gpu.py tests/test_compress_gpu.py -x -q -m gpu 2>&1 | tail -2
for v in 0 1; do echo "REGPANEL=$v"; SB200_QR_REGPANEL=$v python -m pytest tests/test_hss_gpu.py -x -q -m gpu 2>&1 | tail -1; SB200_QR_REGPANEL=$v timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GF/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'qr_ms',round(d['roofline']['kernel_ms'],3), 'resid',d['config']['solve_residual'])"; done
