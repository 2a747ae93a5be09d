#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/r2_accuracy.py 2>&1 | tail -n 4
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 8) | cut -c1-200
