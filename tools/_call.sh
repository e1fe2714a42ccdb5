O=gpurun_out/c19
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cheb.py -m gpu -q -x > $O/cheb_suite.log 2>&1; tail -3 $O/cheb_suite.log | cut -c1-300
for v in "PPS_CHEB_BLOCK=3 PPS_CHEB_F32=1" "PPS_CHEB_BLOCK=4 PPS_CHEB_F32=1" "PPS_CHEB_BLOCK=2 PPS_CHEB_F32=1"; do
  echo "== $v" >> $O/cheb_sweep.log
  env $v timeout 200 python tools/probe.py solve 256 cheb >> $O/cheb_sweep.log 2>&1
  env $v timeout 100 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1 2>&1 | grep -E "finished|SolverInFunction" >> $O/cheb_sweep.log
done
grep -o '"name": "cheb_blocked[^}]*' $O/cheb_sweep.log | cut -c1-200; grep SolverInFunction $O/cheb_sweep.log
