#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2p}
N=${2:-2}
(timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu -x 2>&1 | tail -n 12) | cut -c1-220
for mode in engine py; do
SB200_DIST=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_n${N}_$mode.json 2> gpurun_out/${T}_bench_n${N}_$mode.err
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/${T}_bench_n${N}_$mode.err | tail -n 4 | cut -c1-300
python - <<PY
import json
for ln in open("gpurun_out/${T}_bench_n${N}_$mode.json"):
    if ln.startswith("{"):
        d=json.loads(ln)
        print("N=$N DIST=$mode ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"], "launches", d["gpu_launches"], "dist_parity", d.get("dist_parity"))
PY
done
