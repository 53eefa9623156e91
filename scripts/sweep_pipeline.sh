for cfg in "8 0" "8 6" "8 5" "8 4" "16 5" "4 5" "16 4" "1 0"; do set -- $cfg
python bench.py --batches $1 --paste-ctas $2 --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('batches',d['config']['tile_batches'],'ctas',d['config']['paste_ctas_per_sm'],'ms',round(d['ms_per_step'],2),'e2e_ms',round(d['e2e']['ms_per_step'],2),'k1',round(r['k1_ms_per_step'],2),'k1_alone',round(r['k1_alone_ms'],2))"
done
