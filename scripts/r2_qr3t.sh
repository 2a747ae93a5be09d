#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2e}
timeout 60 python - > gpurun_out/${T}_qrtime.log 2>&1 <<'PY'
import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
for variant in (1,):
    for count in (1, 296, 4096):
        _, _, ms = sb.debug_qr_batch(A, 231, count=count, variant=variant, reps=2)
        print(f"variant {variant} count {count}: {ms:.3f} ms", flush=True)
PY
cat gpurun_out/${T}_qrtime.log | tail -n 60
