#!/bin/bash
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c9; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_join.py -m gpu -q -s > $OUT/pytest_join.log 2>&1; tail -30 $OUT/pytest_join.log | cut -c1-300
timeout 300 python bench.py --no-cpu --no-proxy-leg --steps 6 --no-parity > $OUT/bench.json 2> $OUT/bench.err; grep "closed loop\|value:\|e2e done\|e2e, worker" $OUT/bench.err | cut -c1-500
timeout 300 python bench.py --no-cpu --no-proxy-leg --steps 3 --no-parity --leader-ctas 2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; grep "closed loop\|value:" $OUT/bench_c2.err | cut -c1-400
timeout 300 python bench.py --no-cpu --no-proxy-leg --steps 3 --no-parity --leader-ctas 16 > $OUT/bench_c16.json 2> $OUT/bench_c16.err; grep "closed loop\|value:\|e2e done" $OUT/bench_c16.err | cut -c1-400
