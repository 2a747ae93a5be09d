#!/bin/bash
# full GPU test-suite + optional quick bench
mkdir -p gpurun_out
T=${1:-t}
(time timeout 1200 python -m pytest tests -q -m gpu -x) > gpurun_out/${T}_pytest.log 2>&1; tail -25 gpurun_out/${T}_pytest.log | cut -c1-220
