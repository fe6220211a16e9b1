#!/bin/bash
# last GPU call of round 2: ncu launch list and one full capture of the replica kernel on the bench command, failover re-run
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2cB; mkdir -p $OUT
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > $OUT/ncu_launches.log 2>&1; tail -3 $OUT/launches.csv | cut -c1-300
timeout 150 ncu --set full --clock-control none --import-source on -k apus_replica_kernel -s 2 -c 1 -f -o $OUT/replica_kernel \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > $OUT/ncu_full.log 2>&1; tail -2 $OUT/ncu_full.log | cut -c1-300
timeout 60 ncu -i $OUT/replica_kernel.ncu-rep --page raw --csv > $OUT/replica_kernel_raw.csv 2>/dev/null; wc -c $OUT/replica_kernel_raw.csv
timeout 120 python -m pytest tests/test_gpu_failover.py -m gpu -q -s > $OUT/pytest_failover.log 2>&1; tail -4 $OUT/pytest_failover.log | cut -c1-300
ls -la $OUT
