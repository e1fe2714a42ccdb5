# call 9 (1 GPU): re-run the suite with the rewritten blocked Chebyshev kernel, timing sweeps, bench
O=gpurun_out/c9
mkdir -p $O
export PPS_MARGINS_FILE=$PWD/$O/parity_margins.jsonl
rm -f $PPS_MARGINS_FILE
timeout 1500 python -m pytest tests -m gpu -q -k "not multi" > $O/gpu_suite.log 2>&1
tail -6 $O/gpu_suite.log | cut -c1-300
unset PPS_MARGINS_FILE
for v in "PPS_CHEB_BLOCK=0" "PPS_CHEB_BLOCK=1" "PPS_CHEB_BLOCK=2" "PPS_CHEB_BLOCK=3" "PPS_CHEB_BLOCK=4" "PPS_CHEB_BLOCK=3 PPS_CHEB_F32=1" "PPS_CHEB_BLOCK=4 PPS_CHEB_F32=1" "PPS_CHEB_BLOCK=3 PPS_ZCHUNK_CHEB=32" "PPS_CHEB_BLOCK=3 PPS_ZCHUNK_CHEB=128"; do
  echo "== $v" >> $O/cheb_sweep.log
  env $v timeout 200 python tools/probe.py solve 256 cheb >> $O/cheb_sweep.log 2>&1
  env $v timeout 100 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1 2>&1 | grep -E "finished|SolverInFunction" >> $O/cheb_sweep.log
done
for v in "" "PPS_ZCHUNK_FUSED_P=48" "PPS_ZCHUNK_FUSED_P=85" "PPS_ZCHUNK_FUSED_P=102" "PPS_ZCHUNK_FUSED_P=32" "PPS_FUSE_STAGES_P=4"; do
  echo "== $v" >> $O/sweep.log
  env $v timeout 200 python tools/probe.py solve 512 >> $O/sweep.log 2>&1
done
timeout 500 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err
cut -c1-600 $O/bench.json
PPS_PHASE_TIMERS=1 timeout 100 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1 > $O/driver_phase_report.log 2>&1
timeout 900 python tools/bandwidth_sweep.py > $O/bandwidth_sweep.jsonl 2> $O/bandwidth_sweep.err
tail -3 $O/bandwidth_sweep.jsonl | cut -c1-300
