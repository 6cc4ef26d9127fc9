set -x
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/tests_fullsize_r2f.log 2>&1
grep -E "passed|failed" gpurun_out/tests_fullsize_r2f.log | tail -3
grep -E "^C2 seed|^U: rows|^X: rows|^J: rows|^J_all|^U_best|^X_best|^C4|^C5|near" gpurun_out/tests_fullsize_r2f.log | cut -c1-210
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -4
