#!/bin/bash
# leaf-QR experiment: CTA-cooperative trailing update (debug variant 2 / SB200_QR_VARIANT=5) against the default
mkdir -p gpurun_out
T=${1:-r2u}
(timeout 200 python -m pytest tests/test_qr_kernel_gpu.py -q -x -k "2]" 2>&1 | tail -n 12) | cut -c1-220
timeout 120 python - > gpurun_out/${T}_qrx.log 2>&1 <<'PY'
import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
ref, _, _ = sb.debug_qr_batch(A, 231, count=8, variant=0, reps=1)
for name, v in (("default", 0), ("coop", 2)):
    out, _, _ = sb.debug_qr_batch(A, 231, count=8, variant=v, reps=1)
    R = np.triu(out[:231, :231]); R0 = np.triu(ref[:231, :231])
    err = np.abs(np.abs(R) - np.abs(R0)).max() / np.abs(R0).max()
    line = f"{name:18s} |R| diff {err:.1e}"
    for count in (1, 148, 592, 4096):
        _, _, ms = sb.debug_qr_batch(A, 231, count=count, variant=v, reps=3)
        line += f"  count {count}: {ms:.3f} ms"
    print(line, flush=True)
PY
cat gpurun_out/${T}_qrx.log | cut -c1-200
for env in "SB200_QR_VARIANT=1" "SB200_QR_VARIANT=5"; do
  env $env timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$env', 'ms', round(d['ms_per_step'],3), 'qr_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'resid', d['config']['solve_residual'])"
done 2>&1 | tee gpurun_out/${T}_ab.log
