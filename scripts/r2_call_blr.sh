#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2n}
(time timeout 900 python -m pytest tests/test_blr_gpu.py tests/test_compress_gpu.py tests/test_configs_gpu.py tests/test_parity_at_size_gpu.py -q -m gpu -x 2>&1 | tail -n 8) 2>&1 | cut -c1-200
timeout 900 python bench.py --workload blr --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_blr.json 2> gpurun_out/${T}_bench_blr.err
tail -n 3 gpurun_out/${T}_bench_blr.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_blr.json"))
print("BLR ms", d["value"], "e2e", d["e2e"]["value"], "rank", d["config"]["rank"], "err", d["config"]["solve_rel_err"], "frac", d["roofline"]["frac"])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('HSS ms', d['ms_per_step'], 'compress_s', d['config']['compress_s'], 'err', d['config']['compress_rel_err'])"
