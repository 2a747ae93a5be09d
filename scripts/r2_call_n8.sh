#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2m}
N=${2:-8}
for v in 1 0; do
SB200_GRAPH=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_n${N}_graph$v.json 2> gpurun_out/${T}_bench_n${N}_graph$v.err
tail -n 3 gpurun_out/${T}_bench_n${N}_graph$v.err | cut -c1-300
python - <<PY
import json
for ln in open("gpurun_out/${T}_bench_n${N}_graph$v.json"):
    if ln.startswith("{"):
        d=json.loads(ln)
        print("N=$N GRAPH=$v ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "resid", d["config"]["solve_residual"], "qr_ms", d["roofline"]["kernel_ms"], "dist_parity", d.get("dist_parity"))
PY
done
