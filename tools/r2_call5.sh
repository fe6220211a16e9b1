#!/bin/bash
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c5; mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_v2.py -m gpu -q --maxfail=8 --timeout=200 > $OUT/pytest_v2.log 2>&1; tail -15 $OUT/pytest_v2.log
for i in 1 2 3; do timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sustained or wrap_laps or exact_fit" > $OUT/pytest_sustained_$i.log 2>&1; tail -1 $OUT/pytest_sustained_$i.log; done
timeout 300 python -m pytest tests/test_gpu_join.py -m gpu -q -s > $OUT/pytest_join.log 2>&1; tail -30 $OUT/pytest_join.log
timeout 300 python -m pytest tests/test_gpu_failover.py -m gpu -q -s -k "3" > $OUT/pytest_failover.log 2>&1; tail -4 $OUT/pytest_failover.log
timeout 420 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; grep -v "^$" $OUT/bench.err | grep -v "proxy leg" | tail -12 | cut -c1-600; cut -c1-200 $OUT/bench.json
