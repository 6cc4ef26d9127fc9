set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/tests_v5.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h16_v5.json 2> gpurun_out/bench_h16_v5.err
python tools/ilqr_bench.py --config C1 --B 4096 --maxiter 100 --cpu-states 64 > gpurun_out/ilqr_c1.json 2> gpurun_out/ilqr_c1.err
python tools/ilqr_bench.py --config C1 --B 1 --maxiter 100 --cpu-states 1 --reps 5 > gpurun_out/ilqr_c1_b1.json 2>> gpurun_out/ilqr_c1.err
python tools/ilqr_bench.py --config C2 --B 4096 --maxiter 100 --cpu-states 16 --reps 2 > gpurun_out/ilqr_c2.json 2> gpurun_out/ilqr_c2.err
python tools/ilqr_bench.py --config C1 --B 4096 --maxiter 100 --bilevel > gpurun_out/bilevel_c1.json 2> gpurun_out/bilevel_c1.err
python tools/ilqr_bench.py --config C2 --B 1024 --maxiter 20 --bilevel --reps 2 > gpurun_out/bilevel_c2.json 2> gpurun_out/bilevel_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/launches_ilqr.csv python tools/ilqr_bench.py --config C2 --B 256 --maxiter 3 --cpu-states 0 --reps 1 > gpurun_out/ncu_l_ilqr.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ilqr_kernel -c 1 -o gpurun_out/prof_ilqr_v1 -f python tools/ilqr_bench.py --config C2 --B 256 --maxiter 3 --cpu-states 0 --reps 0 > gpurun_out/ncu_f_ilqr.log 2>&1
tail -3 gpurun_out/tests_v5.log; cat gpurun_out/bench_h16_v5.json gpurun_out/ilqr_c1.json gpurun_out/ilqr_c1_b1.json gpurun_out/ilqr_c2.json gpurun_out/bilevel_c1.json gpurun_out/bilevel_c2.json
