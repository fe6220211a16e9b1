#!/bin/bash
# round 2, GPU call 3 (1 GPU): v2 tests, the autoprune race (repeated), failover, everything else, the bench line.
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c3; mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_v2.py -m gpu -q --maxfail=8 --timeout=150 > $OUT/pytest_v2.log 2>&1; tail -25 $OUT/pytest_v2.log
for i in 1 2 3 4 5 6; do timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sustained or wrap_laps" > $OUT/pytest_sustained_$i.log 2>&1; tail -1 $OUT/pytest_sustained_$i.log; done
timeout 300 python -m pytest tests/test_gpu_failover.py -m gpu -q -x -s > $OUT/pytest_failover.log 2>&1; tail -12 $OUT/pytest_failover.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=5 --deselect tests/test_gpu_v2.py --deselect tests/test_gpu_failover.py > $OUT/pytest_gpu.log 2>&1; tail -6 $OUT/pytest_gpu.log
timeout 420 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -12 $OUT/bench.err; cut -c1-300 $OUT/bench.json
