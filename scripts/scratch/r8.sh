python -m pytest tests/test_gpu_index_build.py -x -q -m gpu 2>&1 | grep -E "^E|^FAILED" | head -6
for R in 2500000 5000000 10000000 1250000; do for C in 1 0; do
ORAG_COSCHEDULE=$C python bench.py --rows $R --no-cpu-baseline --steps 30 2>gpurun_out/st.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print('rows $R cosched $C', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'scan', round(r['launch_ms'],3), round(r['launch_ms_min'],3), 'bm25', round(b['launch_ms'],3), round(b['launch_ms_min'],3), d['clocks']['sm_mhz'])"; tail -1 gpurun_out/st.err
done; done
