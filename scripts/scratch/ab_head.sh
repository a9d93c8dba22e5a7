python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -3; python scripts/scratch/counts.py 1250000; python scripts/scratch/counts.py 10000000
for H in 0 1; do for L in 2 3; do
echo "head=$H lanes=$L"; ORAG_HEAD_STREAM=$H ORAG_LANES=$L python bench.py --rows 1250000 --no-cpu-baseline --steps 40 --timeline gpurun_out/tl3_h${H}_l${L}.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], b.get('launch_ms_in_timed_loop'), d['clocks']['sm_mhz'])"
done; done
for H in 0 1; do
echo "10M head=$H"; ORAG_HEAD_STREAM=$H python bench.py --verify-queries 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], d['clocks']['sm_mhz'], d['verified_against_oracle']['bitwise'])"
done
