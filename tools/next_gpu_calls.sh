#!/bin/bash
# tools/next_gpu_calls.sh -- what to run first when GPU time is available again: the items of DESIGN.md sections 7-9 that
# are written but were not (or only partly) run on hardware when round 2's GPU budget ended.  Each block is one gpurun call.
#
#   tools/gpu.sh --gpus 1 --timeout 600 -- 'bash tools/next_gpu_calls.sh memcached'
#   tools/gpu.sh --gpus 8 --timeout 900 -- 'bash tools/next_gpu_calls.sh fabric'
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/next; mkdir -p $OUT
case "${1:-}" in
  memcached)   # BASELINE config 4 on the GPU engine: the opt-in test, then the launcher with 3 replicas on this box's GPUs
    APUS_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_zz_gpu_memcached_dropin.py -m gpu -q -s > $OUT/pytest_memcached.log 2>&1; tail -5 $OUT/pytest_memcached.log
    timeout 120 bash benchmarks/run_gpu.sh --app=memcached --scount=3 --ccount=16 --rcount=20000 --dsize=1024 > $OUT/run_gpu_memcached.txt 2>&1; tail -4 $OUT/run_gpu_memcached.txt
    ;;
  fabric)      # NVSwitch multicast with more than one follower: parity at 3 and 5 members, then the A/B sweep at 5 and 7
    timeout 200 python -m pytest tests/test_gpu_v2.py -m gpu -q -k multicast > $OUT/pytest_multicast.log 2>&1; tail -4 $OUT/pytest_multicast.log
    for n in 5 7; do
      timeout 150 python tools/sweep_spread.py --replicas $n --sizes 64,1024,4096 --ctas 16,32 --steps 3 --multicast --out $OUT/sweep_${n}_multicast.txt > $OUT/sweep_${n}_mc.log 2>&1
      timeout 150 python tools/sweep_spread.py --replicas $n --sizes 64,1024,4096 --ctas 16,32 --steps 3 --out $OUT/sweep_${n}_plain.txt > $OUT/sweep_${n}_plain.log 2>&1
      tail -4 $OUT/sweep_${n}_multicast.txt $OUT/sweep_${n}_plain.txt
    done
    ;;
  *) echo "usage: $0 memcached|fabric"; exit 1;;
esac
