for L in 2 3 4; do for C in 1 0; do
echo "lanes=$L cosched=$C"; ORAG_LANES=$L ORAG_COSCHEDULE=$C python bench.py --rows 1250000 --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], b.get('launch_ms_in_timed_loop'), d['clocks'])"
done; done
