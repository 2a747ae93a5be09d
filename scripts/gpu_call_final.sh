#!/bin/bash
# end-of-session validation: full GPU suite, smoke, bench (ours + reference arm), launch list, ncu of the streamed solve kernels
mkdir -p gpurun_out
T=r1d
(time timeout 900 python -m pytest tests -q -m gpu -x) > gpurun_out/${T}_pytest.log 2>&1
tail -n 6 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 400 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 1800 gpurun_out/${T}_bench_n1.json
timeout 300 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 900 gpurun_out/${T}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
wc -l gpurun_out/${T}_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ulv_(fwd|bwd)_pipe" -c 22 -o gpurun_out/${T}_ncu_solve -f python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_solve.log 2>&1
ls -la gpurun_out/${T}_ncu_solve.ncu-rep
