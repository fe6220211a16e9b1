#!/bin/bash
# fabric mode on hardware (2 GPUs): multicast parity, then 4 KiB / 64 B A/B with and without multimem.st
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2cC; mkdir -p $OUT
timeout 70 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=60 -k "multicast" > $OUT/pytest_multicast_2gpus.log 2>&1; tail -12 $OUT/pytest_multicast_2gpus.log | cut -c1-300
timeout 40 python tools/sweep_spread.py --replicas 2 --sizes 64,4096 --ctas 16 --steps 3 --multicast --out $OUT/sweep_2gpus_multicast.txt > $OUT/sweep_mc.log 2>&1; tail -3 $OUT/sweep_mc.log | cut -c1-300
timeout 40 python tools/sweep_spread.py --replicas 2 --sizes 64,4096 --ctas 16 --steps 3 --out $OUT/sweep_2gpus_plain.txt > $OUT/sweep_plain.log 2>&1; tail -3 $OUT/sweep_plain.log | cut -c1-300
