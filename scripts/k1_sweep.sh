for eb in 8 12 16; do
  python bench.py --no-cpu-baseline --steps 5 --e2e-batches $eb 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],2), round(r['k1_ms_per_step'],2), round(r['frac'],3), 'e2e', d['config']['e2e_tile_batches'], round(d['e2e']['ms_per_step'],2), d['config']['per_step_ms'])"
done
