python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for R in 1250000 10000000; do
python bench.py --rows $R --verify-queries 8 --steps 30 --timeline gpurun_out/tl9_$R.txt 2>gpurun_out/st.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print($R, d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], d['clocks']['sm_mhz'], d['verified_against_oracle']['bitwise'])"; tail -1 gpurun_out/st.err
done
