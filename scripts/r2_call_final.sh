#!/bin/bash
# full GPU suite, smoke(), then both bench arms and the BLR line
mkdir -p gpurun_out
T=${1:-r2r}
(time timeout 1500 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -n 40) > gpurun_out/${T}_pytest.log 2>&1
tail -n 22 gpurun_out/${T}_pytest.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 3000 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -c 1500 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --workload blr > gpurun_out/${T}_bench_blr.json 2> gpurun_out/${T}_bench_blr.err; tail -c 1500 gpurun_out/${T}_bench_blr.json
