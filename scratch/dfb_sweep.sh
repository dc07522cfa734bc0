for c in 2 3 4; do
  GIGS_DFB=$c python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); sm=d['stage_ms']
        print('dfb $c', round(d['ms_per_step'],4), {k:round(v,4) for k,v in sm.items() if k in ('deferred_backward','deferred_shade','deferred_loss','blend_backward','blend_forward')})
"
done
