python -m pytest tests/test_gpu_index_build.py -x -q -m gpu 2>&1 | tail -1
for C in 1 0; do
ORAG_COSCHEDULE=$C python bench.py --rows 1250000 --no-cpu-baseline --steps 40 2>gpurun_out/st.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print('cosched $C 1.25M', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], d['clocks']['sm_mhz'])"; tail -1 gpurun_out/st.err
done
python bench.py > gpurun_out/r02_bench_config3_d.json 2>gpurun_out/st.err; tail -1 gpurun_out/st.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_config3_d.json').read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print('10M', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r['frac'], 'bm25', b['launch_ms'], b['launch_ms_min'], d['clocks'], d['verified_against_oracle']['bitwise'])"
