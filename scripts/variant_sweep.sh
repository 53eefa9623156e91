#!/bin/bash
# usage: scripts/variant_sweep.sh <suffix> ...   -- benches deepemia_b200/libemia_<suffix>.so one after the other (stage breakdown)
cp deepemia_b200/libemia.so /tmp/libemia_keep.so
for c in "$@"; do
  cp deepemia_b200/libemia_$c.so deepemia_b200/libemia.so
  echo "variant $c"
  python bench.py --no-cpu-baseline --breakdown --steps 3 ${SWEEP_ARGS} 2>&1 >/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k: round(v,3) for k,v in d['stage_ms'].items()}, round(d['sum_ms'],2))"
done
cp /tmp/libemia_keep.so deepemia_b200/libemia.so
