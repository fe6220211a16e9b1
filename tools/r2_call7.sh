#!/bin/bash
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c7; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=200 -k "slow_follower" > $OUT/pytest_slow.log 2>&1; grep -n "AssertionError\|leader offsets\|apply offsets\|leader stats\|passed\|failed" $OUT/pytest_slow.log | cut -c1-1500 | head -20
timeout 300 python -m pytest tests/test_gpu_join.py -m gpu -q -s > $OUT/pytest_join.log 2>&1; tail -40 $OUT/pytest_join.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=200 -k "express or heartbeat or term or stop_while" > $OUT/pytest_v2.log 2>&1; tail -3 $OUT/pytest_v2.log
timeout 300 python bench.py --no-cpu --no-proxy-leg --steps 6 --no-parity > $OUT/bench.json 2> $OUT/bench.err; grep "closed loop\|value:\|e2e done" $OUT/bench.err | cut -c1-500
timeout 300 python bench.py --no-cpu --no-proxy-leg --steps 4 --no-parity --profile-latency > $OUT/bench_prof.json 2> $OUT/bench_prof.err; grep "closed loop" $OUT/bench_prof.err | cut -c1-500
