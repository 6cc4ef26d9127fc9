set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/tests_fullsize_r2d.log
tail -60 gpurun_out/tests_fullsize_r2d.log
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -6 > gpurun_out/tests_r2d.log
cat gpurun_out/tests_r2d.log
