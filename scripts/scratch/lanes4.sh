for L in 4 2; do
echo "lanes $L"
ORAG_LANES=$L python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --rows 2500000 --no-cpu-baseline --steps 40 2>gpurun_out/n2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
ORAG_LANES=$L python bench.py --rows 1250000 --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=1', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
