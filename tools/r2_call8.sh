#!/bin/bash
# round 2, 8-GPU call: cross-GPU byte parity of the gpu test-suite, the north_star sweep with one replica per GPU and the
# leader GPU's NVLink counters, closed-loop latency at 5 replica GPUs, torchrun N=8 with parity, failover / join / redis
# with one GPU per replica.  Most valuable first: a cut-off call still leaves the essentials.
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c8; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nvidia-smi nvlink -gt d -i 0 >> $OUT/gpus.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_v2.py -m gpu -q --maxfail=6 --timeout=120 > $OUT/pytest_gpu_8gpus.log 2>&1; tail -4 $OUT/pytest_gpu_8gpus.log
timeout 200 python tools/sweep_spread.py --ctas 16 --steps 3 --out $OUT/sweep_ctas16.txt > $OUT/sweep16.log 2>&1; tail -14 $OUT/sweep16.log
timeout 120 python tools/sweep_spread.py --sizes 1024,4096 --replicas 7 --ctas 32,64 --steps 3 --out $OUT/sweep_ctas32_64.txt > $OUT/sweep64.log 2>&1; tail -6 $OUT/sweep64.log
timeout 200 python bench.py --spread --replicas 5 --steps 8 --no-cpu > $OUT/bench_spread5.json 2> $OUT/bench_spread5.err; grep -v "^$" $OUT/bench_spread5.err | tail -8 | cut -c1-600
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu > $OUT/bench_8gpu_torchrun.json 2> $OUT/bench_8gpu.err; grep "parity\|value:" $OUT/bench_8gpu.err | cut -c1-400; cut -c1-300 $OUT/bench_8gpu_torchrun.json
timeout 150 python bench.py --failover --failover-trials 2 > $OUT/failover.json 2> $OUT/failover.err; cut -c1-1500 $OUT/failover.json
timeout 150 python -m pytest tests/test_gpu_failover.py tests/test_gpu_join.py -m gpu -q -s > $OUT/pytest_failover_join_8gpus.log 2>&1; tail -5 $OUT/pytest_failover_join_8gpus.log
timeout 100 bash benchmarks/run_gpu.sh --app=redis --scount=5 --ccount=16 --rcount=200000 > $OUT/redis_5gpus.txt 2>&1; tail -6 $OUT/redis_5gpus.txt
timeout 100 bash benchmarks/run_gpu.sh --app=redis --scount=5 --ccount=16 --rcount=100000 --kill-leader --port=9888 > $OUT/redis_5gpus_kill_leader.txt 2>&1; tail -8 $OUT/redis_5gpus_kill_leader.txt
ls -la $OUT
