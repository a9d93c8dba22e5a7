for P in 0 8; do for L in 2 3; do
echo "probe=$P lanes=$L"; ORAG_PROBE_QSHARD=$P ORAG_HEAD_STREAM=0 ORAG_LANES=$L python bench.py --rows 1250000 --no-cpu-baseline --steps 40 --timeline gpurun_out/tl4_p${P}_l${L}.txt 2>gpurun_out/qs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print(d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], b.get('launch_ms_in_timed_loop'), d['clocks']['sm_mhz'])"; tail -2 gpurun_out/qs.err
done; done
