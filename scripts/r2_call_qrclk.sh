#!/bin/bash
# phase clocks (-DSB200_QR_TIMING build, scripts/libsb200_timing.so) of a leaf QR variant: 1 leaf alone and 4 per SM
mkdir -p gpurun_out
T=${1:-r2w}; V=${2:-2}
SB200_LIB=$PWD/scripts/libsb200_timing.so timeout 60 python - $V > gpurun_out/${T}_qrclk.log 2>&1 <<'PY'
import sys, numpy as np, strumpack_b200 as sb
v = int(sys.argv[1])
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
for count in (1, 592):
    print("count", count, flush=True)
    sb.debug_qr_batch(A, 231, count=count, variant=v, reps=1)
PY
grep -E "count|block 0 warp [03]" gpurun_out/${T}_qrclk.log | cut -c1-330
