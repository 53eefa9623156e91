python -m pytest tests -m gpu -x -q -k "paste" 2>&1 | tail -2
for v in 2 3; do
  python bench.py --no-cpu-baseline --steps 5 --variant $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('variant',d['config']['paste_variant'], round(d['ms_per_step'],2), 'k1', round(r['k1_ms_per_step'],2), round(r['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e_fp16_heads']['ms_per_step'],2), d['e2e_fp16_heads']['identical_results_to_f32'])"
done
