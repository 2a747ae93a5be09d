#!/bin/bash
mkdir -p gpurun_out
for so in "$@"; do
  cp $so strumpack_b200/libstrumpack_b200.so
  echo "== $so"
  timeout 60 python - 2>&1 <<'PY' | grep -v "warp [123]:" | tail -n 12
import numpy as np, strumpack_b200 as sb
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((256, 281)))
for count in (1, 4096):
    _, _, ms = sb.debug_qr_batch(A, 231, count=count, variant=1, reps=2)
    print(f"variant 1 count {count}: {ms:.3f} ms", flush=True)
PY
done
