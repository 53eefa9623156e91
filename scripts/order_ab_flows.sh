# A/B of the length-sorted K5 work order on the batched flows
for o in 0 1 0 1; do
  EMIA_MEASURE_ORDER=$o python bench.py --flows-only config3a,config3b,config4 --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null > gpurun_out/ab_$o.json
  python - "$o" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
print("order", sys.argv[1], {k: round(v["ms_per_step"], 3) for k, v in d.items()})
PY
done
