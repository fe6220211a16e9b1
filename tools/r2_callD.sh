#!/bin/bash
# fabric mode, ragged 1000 B entries (byte-granular tile edges next to multimem.st chunks), 2 GPUs
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2cD; mkdir -p $OUT
timeout 60 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=50 -k "multicast and 1000" > $OUT/pytest_multicast_1000B_2gpus.log 2>&1; tail -6 $OUT/pytest_multicast_1000B_2gpus.log | cut -c1-300
