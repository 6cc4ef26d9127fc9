set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/tests_r2h.log; cat gpurun_out/tests_r2h.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -2 gpurun_out/bench_r2h.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r2h.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'path', 'clamped_ranks', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'traffic', 'frac_of_ceiling')})
print('e2e', d['e2e']['value'])
print('parity U', d['parity']['U'])
for k, v in d['configs'].items(): print(k, v['ms_per_step'], v['path'], round(v['tflops_per_gpu'], 1), v.get('acting_latency_get_optimal_action'))
P
