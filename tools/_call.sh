# call 12 (1 GPU): suite + smoke + bench with the final defaults; ncu launch list of the bench
O=gpurun_out/c12
mkdir -p $O
export PPS_MARGINS_FILE=$PWD/$O/parity_margins.jsonl
rm -f $PPS_MARGINS_FILE
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1
tail -6 $O/gpu_suite.log | cut -c1-300
unset PPS_MARGINS_FILE
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 500 python bench.py > $O/bench.json 2> $O/bench.err
cut -c1-400 $O/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
cut -c1-1500 $O/bench_reference.json
for v in "PPS_GRAPH=0 PPS_BATCH_GHOSTS=0" "PPS_GRAPH=0" ""; do echo "== $v" >> $O/default_problem.log; env $v timeout 100 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1 2>&1 | grep -E "finished|SolverInFunction" >> $O/default_problem.log; done
cat $O/default_problem.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file $O/launches_bench512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-gate > $O/bench_under_ncu.log 2>&1
