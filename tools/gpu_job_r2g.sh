set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/tests_multi_r2g.log; cat gpurun_out/tests_multi_r2g.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2g_${N}gpu.json 2> gpurun_out/bench_r2g_${N}gpu.err
tail -3 gpurun_out/bench_r2g_${N}gpu.err
[ "$N" = "2" ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_r2g_${N}gpu.json 2> gpurun_out/bench_ref_r2g_${N}gpu.err
python - <<P
import json
d = json.loads(open('gpurun_out/bench_r2g_${N}gpu.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'path', 'clamped_ranks')})
print('e2e', d['e2e'])
print('shard_parity', d.get('shard_parity'))
print('critic', d.get('critic'))
print('C5', d.get('configs', {}).get('C5'))
P
