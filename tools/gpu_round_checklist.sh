#!/bin/bash
# tools/gpu_round_checklist.sh -- one gpurun call that collects everything a round needs from a 1-GPU B200 box
# (run as:  gpurun --timeout 900 -- 'bash tools/gpu_round_checklist.sh' ; outputs land in gpurun_out/round/).
# Order = most valuable first, so a cut-off call still leaves the essentials.
set -u
OUT=gpurun_out/round; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 300 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; cut -c1-300 $OUT/bench_1gpu.json
timeout 300 python bench.py --impl reference > $OUT/bench_1gpu_reference_arm.json 2> $OUT/bench_ref.err
# launch list of the same command (per-launch times are cold-cache and serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_launches.log 2>&1
# one full capture of the replica kernel (skip the warm-up launch)
timeout 600 ncu --set full --clock-control none --import-source on -s 2 -c 1 -o $OUT/replica_kernel \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
# NVLink counters exist only with >= 2 GPUs: list what this ncu offers so the 2-GPU call can name them
ncu --query-metrics 2>/dev/null | grep -i -E "nvl|nvlink" | head -60 > $OUT/nvlink_metric_names.txt
ls -la $OUT
