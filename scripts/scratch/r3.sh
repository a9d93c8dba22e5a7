python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -2
python scripts/scratch/counts.py 1250000
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_dist.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --rows 2500000 --no-cpu-baseline --steps 40 --timeline gpurun_out/tl8_n2.txt 2>gpurun_out/n2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print('N=2 1.25M/GPU', d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], d['clocks']['sm_mhz'], d.get('per_rank_kernel_ms'))"
for R in 1250000 10000000; do
python bench.py --rows $R --verify-queries 8 --steps 30 --timeline gpurun_out/tl8_$R.txt 2>gpurun_out/st.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['roofline_bm25']
print($R, d['ms_per_step'], d['e2e']['ms_per_step'], 'scan', r['launch_ms'], r['launch_ms_min'], r.get('launch_ms_in_timed_loop'), 'bm25', b['launch_ms'], b['launch_ms_min'], d['clocks']['sm_mhz'], d['verified_against_oracle'])"; tail -1 gpurun_out/st.err
done
