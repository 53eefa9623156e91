for cfg in "1 0" "4 0" "8 0" "8 4" "16 0" "16 4"; do set -- $cfg
python bench.py --batches $1 --paste-ctas $2 --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('batches',d['config']['tile_batches'],'ctas',d['config']['paste_ctas_per_sm'],'ms',round(d['ms_per_step'],2),'k1',round(r['k1_ms_per_step'],2), d['config']['per_step_ms'])"
done
