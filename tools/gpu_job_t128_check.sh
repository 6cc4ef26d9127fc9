# one-GPU check of plan_t128 after a change: timed harness (wait time-outs instead of hangs), plain harness, parity tests of the path
set -x
mkdir -p gpurun_out
L=gpurun_out/t128_bench_check.log
timeout 60 ./build/t128_bench_timed 128 4 2 200 1 > $L 2>&1
timeout 60 ./build/t128_bench_timed 4096 32 20 200 2 >> $L 2>&1
timeout 60 ./build/t128_bench 4096 32 20 200 3 >> $L 2>&1
timeout 60 ./build/t128_bench 18944 32 20 200 3 >> $L 2>&1
cat $L
timeout 600 python -m pytest tests/test_gpu_planner.py -m gpu -q -x -k "t128" 2>&1 | tail -15 > gpurun_out/tests_t128_check.log
cat gpurun_out/tests_t128_check.log
