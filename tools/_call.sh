# call 8 (1 GPU): full GPU suite (blocked / fp32 / global Chebyshev new), bench with the wave-aware z-chunks, tuning experiments
O=gpurun_out/c8
mkdir -p $O
export PPS_MARGINS_FILE=$PWD/$O/parity_margins.jsonl
rm -f $PPS_MARGINS_FILE
timeout 1500 python -m pytest tests -m gpu -q -k "not multi" > $O/gpu_suite.log 2>&1
tail -8 $O/gpu_suite.log | cut -c1-300
unset PPS_MARGINS_FILE
timeout 500 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err
cut -c1-2500 $O/bench.json
tail -2 $O/bench.err | cut -c1-300
# tuning: per-kernel timings of a full 512^3 solve
for v in "" "PPS_TMA_L2PROMO=2" "PPS_TMA_L2PROMO=0" "PPS_FUSE_STAGES_P=3" "PPS_ZCHUNK_STENCIL=32" "PPS_ZCHUNK_STENCIL=43" "PPS_ZCHUNK_STENCIL=128"; do
  echo "== $v" >> $O/sweep.log
  env $v timeout 200 python tools/probe.py solve 512 >> $O/sweep.log 2>&1
done
# Chebyshev-preconditioned solves: per-sweep vs blocked depths vs fp32, 256^3 one block and the shipped default problem
for v in "PPS_CHEB_BLOCK=0" "PPS_CHEB_BLOCK=1" "PPS_CHEB_BLOCK=2" "PPS_CHEB_BLOCK=3" "PPS_CHEB_BLOCK=4" "PPS_CHEB_BLOCK=3 PPS_CHEB_F32=1" "PPS_CHEB_BLOCK=4 PPS_CHEB_F32=1"; do
  echo "== $v" >> $O/cheb_sweep.log
  env $v timeout 200 python tools/probe.py solve 256 cheb >> $O/cheb_sweep.log 2>&1
  env $v timeout 100 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1 2>&1 | grep -E "finished|SolverInFunction" >> $O/cheb_sweep.log
done
PPS_PHASE_TIMERS=1 PPS_CHEB_BLOCK=3 timeout 100 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1 > $O/driver_phase_report.log 2>&1
grep -c . $O/sweep.log $O/cheb_sweep.log
