#!/bin/bash
# source-level ncu capture of the leaf QR kernel (2 full waves of 256 x 281 leaves through the test hook)
mkdir -p gpurun_out
T=${1:-r2t}
cat > gpurun_out/qrrun.py <<'PY'
import sys; sys.path.insert(0, "."); import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
import os; sb.debug_qr_batch(A, 231, count=1184, variant=int(os.environ.get("QRV", "0")), reps=1)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ulv_qr_kernel -c 1 -o gpurun_out/${T}_qr -f python gpurun_out/qrrun.py > gpurun_out/${T}_ncu.log 2>&1
tail -n 3 gpurun_out/${T}_ncu.log | cut -c1-200
ncu -i gpurun_out/${T}_qr.ncu-rep --page raw --csv > gpurun_out/${T}_qr_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_qr.ncu-rep --page source --csv > gpurun_out/${T}_qr_source.csv 2>/dev/null
ncu -i gpurun_out/${T}_qr.ncu-rep --page source --csv --print-source cuda > gpurun_out/${T}_qr_source_cuda.csv 2>/dev/null
ls -la gpurun_out/${T}_qr* | awk '{print $5, $9}'
rm -f gpurun_out/${T}_qr.ncu-rep
head -c 1500 gpurun_out/${T}_qr_source.csv
