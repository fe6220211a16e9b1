#!/bin/bash
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2cA; mkdir -p $OUT
timeout 280 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $OUT/pytest_gpu_all.log 2>&1; tail -25 $OUT/pytest_gpu_all.log | cut -c1-400
timeout 150 python bench.py --no-cpu --no-proxy-leg --steps 8 > $OUT/bench.json 2> $OUT/bench.err; grep "closed loop\|value:\|e2e done\|e2e, worker\|parity" $OUT/bench.err | cut -c1-420
