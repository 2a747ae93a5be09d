#!/bin/bash
# leaf-QR A/B: kernel tests, isolated kernel times (test hook), bench step
mkdir -p gpurun_out
T=${1:-r2u}
(timeout 200 python -m pytest tests/test_qr_kernel_gpu.py -q -x -k "0]" 2>&1 | tail -n 5) | cut -c1-220
timeout 120 python - > gpurun_out/${T}_qrx.log 2>&1 <<'PY'
import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
line = "default"
for count in (1, 148, 592, 4096):
    _, _, ms = sb.debug_qr_batch(A, 231, count=count, variant=0, reps=3)
    line += f"  count {count}: {ms:.3f} ms"
print(line, flush=True)
PY
cat gpurun_out/${T}_qrx.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],3), 'qr_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'resid', d['config']['solve_residual'])" 2>&1 | tee gpurun_out/${T}_ab.log
