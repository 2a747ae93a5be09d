#!/bin/bash
# ulv_qr3 bring-up: kernel tests with a hang guard, then kernel timing old vs new
mkdir -p gpurun_out
T=${1:-r2b}
(timeout 150 python -m pytest tests/test_qr_kernel_gpu.py -q -x 2>&1 | grep -v "qr3 timing" | tail -n 30) > gpurun_out/${T}_qrtest.log 2>&1
cut -c1-250 gpurun_out/${T}_qrtest.log | tail -n 15
timeout 60 python - > gpurun_out/${T}_qrtime.log 2>&1 <<'PY'
import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
for variant in (1, 0):
    for count in (1, 296, 4096):
        _, _, ms = sb.debug_qr_batch(A, 231, count=count, variant=variant, reps=2)
        print(f"variant {variant} count {count}: {ms:.3f} ms", flush=True)
PY
grep -v "warp [123]:" gpurun_out/${T}_qrtime.log | tail -n 40
