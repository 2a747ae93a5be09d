#!/bin/bash
# full GPU suite (default kernels), then the HSS suites again with the optional qr3 kernel
mkdir -p gpurun_out
T=${1:-r2h}
(time timeout 1200 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -n 40) > gpurun_out/${T}_pytest.log 2>&1
tail -n 25 gpurun_out/${T}_pytest.log | cut -c1-220
(time SB200_QR3=1 timeout 600 python -m pytest tests/test_hss_gpu.py tests/test_schur_gpu.py tests/test_qr_kernel_gpu.py -q -m gpu 2>&1 | tail -n 15) > gpurun_out/${T}_pytest_qr3.log 2>&1
tail -n 8 gpurun_out/${T}_pytest_qr3.log | cut -c1-220
