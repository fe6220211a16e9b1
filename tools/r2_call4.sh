#!/bin/bash
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c4; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_v2.py -m gpu -q --maxfail=8 --timeout=200 -k "slow_follower or stop_while or heartbeat or term_fence or express" > $OUT/pytest_v2.log 2>&1; tail -15 $OUT/pytest_v2.log
timeout 300 python -m pytest tests/test_gpu_failover.py -m gpu -q -s > $OUT/pytest_failover.log 2>&1; tail -12 $OUT/pytest_failover.log
timeout 300 python -m pytest tests/test_gpu_join.py -m gpu -q -s > $OUT/pytest_join.log 2>&1; tail -25 $OUT/pytest_join.log
timeout 420 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; grep -v "^$" $OUT/bench.err | tail -14; cut -c1-200 $OUT/bench.json
