#!/bin/bash
# round 2, GPU call 1 (gpurun --gpus 2): fabric probe, latency micro-benchmarks, the gpu test-suite with replicas on
# two GPUs (every "peer" store crosses NVLink), NVLink byte counters of the replica kernel.
set -u
export APUS_NO_BUILD=1
OUT=gpurun_out/r2c1; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
for m in info st mc bulk mcbulk; do timeout 90 tools/probe_fabric $m > $OUT/probe_$m.txt 2>&1; echo "probe $m rc=$?"; done
timeout 120 tools/ubench > $OUT/ubench_2gpu.txt 2>&1; tail -4 $OUT/ubench_2gpu.txt
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_2gpus.log 2>&1; tail -3 $OUT/pytest_gpu_2gpus.log
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink" > $OUT/nvlink_metric_names.txt
M=$(awk '{print $1}' $OUT/nvlink_metric_names.txt | grep -i -E "bytes" | head -16 | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
if [ -n "$M" ]; then
  timeout 300 ncu --metrics "$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum" --clock-control none -s 2 -c 4 --csv \
      --log-file $OUT/nvlink_64B.csv python bench.py --spread --replicas 2 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_nvlink_64B.log 2>&1
  timeout 300 ncu --metrics "$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum" --clock-control none -s 2 -c 4 --csv \
      --log-file $OUT/nvlink_4K.csv python bench.py --spread --replicas 2 --payload 4096 --batch 4096 --leader-ctas 16 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_nvlink_4K.log 2>&1
fi
ls -la $OUT
