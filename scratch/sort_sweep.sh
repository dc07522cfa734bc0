for c in 0 1 2 3 4 5; do
  GIGS_RS_CFG=$c python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); sm=d['stage_ms']
        print('cfg $c', round(d['ms_per_step'],4), {k:round(v,4) for k,v in sm.items() if k in ('depth_sort','emit_keys','radix_sort','radix_sort_pass')})
"
done
