# round-2 evidence run (one GPU): full-size parity tests, the bench line of both arms, launch list, ncu captures
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/tests_fullsize_r2e.log 2>&1
grep -E "passed|failed" gpurun_out/tests_fullsize_r2e.log | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
tail -3 gpurun_out/bench_r2e.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2e.json 2> gpurun_out/bench_ref_r2e.err
python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_r2e.json 2> gpurun_out/bench_c5_r2e.err
python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_r2e.json 2> gpurun_out/bench_c3_r2e.err
# launch list of the default bench command (same command as above, shorter)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_l_r2e.log 2>&1
# full captures of the two planner kernels
ncu --set full --clock-control none --import-source on -k regex:plan_h16 -c 1 -o gpurun_out/prof_h16_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_f_h16_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_t128 -c 1 -o gpurun_out/prof_t128_r2 -f python bench.py --workload C5 --batch 18944 --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_f_t128_r2.log 2>&1
tail -2 gpurun_out/ncu_f_h16_r2.log gpurun_out/ncu_f_t128_r2.log
ls -la gpurun_out/*.ncu-rep | tail -3
