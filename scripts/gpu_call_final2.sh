#!/bin/bash
# last validation of the round: full GPU suite + smoke + default bench
mkdir -p gpurun_out
T=r1e
(time timeout 900 python -m pytest tests -q -m gpu -x) > gpurun_out/${T}_pytest.log 2>&1
tail -n 6 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 400 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 600 gpurun_out/${T}_bench_n1.json
