#!/bin/bash
# round 2, GPU call 2 (1 GPU): the whole gpu test-suite on the v2 kernels, then the default bench line.
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c2; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 -x --timeout=120 tests/test_gpu_v2.py > $OUT/pytest_v2.log 2>&1; tail -15 $OUT/pytest_v2.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=5 --deselect tests/test_gpu_v2.py > $OUT/pytest_gpu.log 2>&1; tail -6 $OUT/pytest_gpu.log
timeout 420 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -12 $OUT/bench.err; cut -c1-600 $OUT/bench.json
