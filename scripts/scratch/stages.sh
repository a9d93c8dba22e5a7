python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for S in 6 5; do for R in 1250000 10000000; do
echo "stages=$S rows=$R"; ORAG_SCAN_STAGES=$S ORAG_HEAD_STREAM=0 python bench.py --rows $R --no-cpu-baseline --steps 30 --timeline gpurun_out/tl7_s${S}_${R}.txt 2>gpurun_out/st.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], b.get('launch_ms_in_timed_loop'), d['clocks']['sm_mhz'])"; tail -1 gpurun_out/st.err
done; done
