#!/bin/bash
# tools/gpu_round_checklist_2gpu.sh -- NVLink evidence on a 2-GPU box (gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_round_checklist_2gpu.sh').
# One process, replica r on GPU r (bench.py --spread): the leader's peer stores cross NVLink, so the nvlink byte
# counters north_star asks for become observable.  ncu replays every kernel: this is evidence, never a bench value.
set -u
OUT=gpurun_out/round2gpu; mkdir -p $OUT
tools/ubench > $OUT/ubench_2gpu.txt 2>&1 || (make -C tools > /dev/null 2>&1 && tools/ubench > $OUT/ubench_2gpu.txt 2>&1)
tail -12 $OUT/ubench_2gpu.txt
timeout 200 python bench.py --spread --replicas 2 --no-cpu > $OUT/bench_spread_2.json 2> $OUT/bench_spread_2.err
timeout 200 python bench.py --spread --replicas 2 --payload 4096 --batch 4096 --leader-ctas 16 --no-cpu --no-e2e > $OUT/bench_spread_2_4k.json 2>> $OUT/bench_spread_2.err
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink" > $OUT/nvlink_metric_names.txt
M=$(awk '{print $1}' $OUT/nvlink_metric_names.txt | grep -i -E "bytes" | head -12 | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
if [ -n "$M" ]; then
  timeout 400 ncu --metrics "$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum" --clock-control none -s 2 -c 4 --csv \
      --log-file $OUT/nvlink_64B.csv python bench.py --spread --replicas 2 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_nvlink_64B.log 2>&1
  timeout 400 ncu --metrics "$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum" --clock-control none -s 2 -c 4 --csv \
      --log-file $OUT/nvlink_4K.csv python bench.py --spread --replicas 2 --payload 4096 --batch 4096 --leader-ctas 16 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_nvlink_4K.log 2>&1
fi
ls -la $OUT
