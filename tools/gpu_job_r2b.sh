set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./build/t128_bench 18944 32 20 200 3 > gpurun_out/t128_bench_r2b.log 2>&1
./build/t128_bench_timed 18944 32 20 200 2 >> gpurun_out/t128_bench_r2b.log 2>&1
./build/t128_bench 4096 32 20 200 3 >> gpurun_out/t128_bench_r2b.log 2>&1
cat gpurun_out/t128_bench_r2b.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/tests_r2b.log
cat gpurun_out/tests_r2b.log
