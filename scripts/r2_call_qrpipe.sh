#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2o}
(timeout 300 python -m pytest tests/test_qr_kernel_gpu.py tests/test_hss_gpu.py tests/test_schur_gpu.py -q -m gpu -x 2>&1 | tail -n 5) | cut -c1-200
timeout 60 python - <<'PY'
import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
for count in (296, 592, 4096):
    _, _, ms = sb.debug_qr_batch(A, 231, count=count, variant=0, reps=3)
    print(f"variant 0 count {count}: {ms:.3f} ms", flush=True)
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('HSS ms', d['ms_per_step'], 'qr_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'resid', d['config']['solve_residual'])"
