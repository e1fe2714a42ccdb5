# call 6 (1 GPU): full GPU suite with the generalised fused schedule, bench, ncu captures, fused-kernel sweeps
O=gpurun_out/c6
mkdir -p $O
export PPS_MARGINS_FILE=$PWD/$O/parity_margins.jsonl
rm -f $PPS_MARGINS_FILE
timeout 1200 python -m pytest tests -m gpu -q -k "not multi" > $O/gpu_suite.log 2>&1
tail -5 $O/gpu_suite.log
unset PPS_MARGINS_FILE
timeout 500 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err
cut -c1-1200 $O/bench.json
tail -3 $O/bench.err
# launch list (shares) of the bench command: 300 launches from the middle of a solve
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file $O/launches_bench512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-gate > $O/bench_under_ncu.log 2>&1
# --set full: one launch each of the three kernels of the fused iteration + the split stencil_dot2
for k in PreSUpdate PrePUpdate OpXRUpdateS; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 10 -c 1 -f -o $O/prof_r02_512_$k python tools/probe.py iters 512 30 > $O/ncu_$k.log 2>&1
done
PPS_FUSE=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiStoreDot2Self -s 10 -c 1 -f -o $O/prof_r02_512_EpiStoreDot2Self python tools/probe.py iters 512 30 > $O/ncu_dot2self.log 2>&1
# sweeps: per-kernel CUDA-event timings of a full 512^3 solve
for v in "" "PPS_ZCHUNK_STENCIL=16" "PPS_ZCHUNK_STENCIL=64" "PPS_FUSE_STAGES_S=4" "PPS_FUSE_STAGES_S=3" "PPS_FUSE_STAGES_P=3" "PPS_FUSE=1" "PPS_FUSE=1 PPS_ZCHUNK_STENCIL=16"; do
  echo "== $v" >> $O/sweep.log
  env $v timeout 200 python tools/probe.py solve 512 >> $O/sweep.log 2>&1
done
grep -c . $O/sweep.log
