set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/tests_v7.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v7.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h16_v7.json 2> gpurun_out/bench_h16_v7.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v7.json 2> gpurun_out/bench_ref_v7.err
python tools/ilqr_bench.py --config C1 --B 4096 --maxiter 100 --cpu-states 0 > gpurun_out/ilqr_c1_v7.json 2>&1
python tools/ilqr_bench.py --config C2 --B 4096 --maxiter 100 --cpu-states 0 --reps 2 > gpurun_out/ilqr_c2_v7.json 2>&1
tail -3 gpurun_out/tests_v7.log; tail -2 gpurun_out/smoke_v7.log; cut -c1-300 gpurun_out/bench_h16_v7.json; cut -c1-200 gpurun_out/bench_ref_v7.json; cut -c1-160 gpurun_out/ilqr_c1_v7.json gpurun_out/ilqr_c2_v7.json
