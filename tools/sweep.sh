#!/bin/bash
# tools/sweep.sh [extra bench.py flags] -- committed ops/s over request size x group size
# (north_star sweep: 64 B - 4 KB requests at 1/3/5/7 replicas), value mode only.
cd "$(dirname "$0")/.."
echo "replicas payload batch ctas ops_per_s alg_GBps ms_per_step"
for n in 1 3 5 7; do
  for spec in "64 65536" "256 32768" "1024 16384" "4096 4096"; do
    set -- $spec; L=$1; B=$2
    out=$(timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --replicas $n --payload $L --batch $B --leader-ctas ${CTAS:-8} "${@:3}" $EXTRA 2>/dev/null)
    python - "$n" "$L" "$B" "${CTAS:-8}" <<PY
import json,sys
try:
    d=json.loads('''$out''')
    n,L,B,c=map(int,sys.argv[1:5])
    print(n,L,B,c,int(d["value"]),round(d["value"]*(n-1)*(64+L)/1e9,2),d["ms_per_step"])
except Exception as e:
    print(sys.argv[1:5],"FAILED",e)
PY
  done
done
