set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/tests_r2a.log
tail -8 gpurun_out/tests_r2a.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -3 gpurun_out/bench_r2a.err
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
    for k in ('value', 'ms_per_step', 'path', 'clamped_ranks'):
        print(k, d.get(k))
    print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'frac_of_ceiling', 'f16_mma_peak_tflops')})
    print('e2e', d['e2e'])
    print('parity', json.dumps(d.get('parity')))
    print('configs', json.dumps(d.get('configs')))
    print('critic', json.dumps(d.get('critic')))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('parse failed', e)
P
