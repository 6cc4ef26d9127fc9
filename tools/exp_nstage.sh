python -m pytest tests/test_gpu_ilqr.py -k edge_shapes -x -q -s 2>&1 | grep -v Warning | tail -40
for N in 3 4 5; do
  if [ $N = 3 ]; then L=gan_mpc_b200/libgmpc.so; else L=build/libgmpc_n$N.so; fi
  echo "NSTAGE=$N"
  GMPC_LIB_PATH=$PWD/$L python tools/ilqr_bench.py --config C2 --B 256 --maxiter 20 --cpu-states 0 --reps 2 | cut -c1-140
  GMPC_LIB_PATH=$PWD/$L python tools/ilqr_bench.py --config C1 --B 1 --maxiter 100 --cpu-states 0 --reps 3 | cut -c1-140
  GMPC_LIB_PATH=$PWD/$L python bench.py --path ffma --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
done
