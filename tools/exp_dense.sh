for L in gan_mpc_b200/libgmpc.so build/libgmpc_dense.so; do
  echo "LIB=$L"
  GMPC_LIB_PATH=$PWD/$L python tools/ilqr_bench.py --config C2 --B 256 --maxiter 20 --cpu-states 0 --reps 2 | cut -c1-140
  GMPC_LIB_PATH=$PWD/$L python bench.py --path ffma --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
done
GMPC_LIB_PATH=$PWD/build/libgmpc_dense.so python -m pytest tests/test_gpu_planner.py tests/test_gpu_ilqr.py tests/test_gpu_dynfit.py tests/test_gpu_bilevel.py -q -k "ffma or ilqr or dynfit or bilevel" 2>&1 | tail -4
