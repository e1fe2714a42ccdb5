# call 13 (1 GPU): final defaults (fused_s on 64x16 tiles): suite, bench (driver's arguments), profile summaries
O=gpurun_out/c13
mkdir -p $O
export PPS_MARGINS_FILE=$PWD/$O/parity_margins.jsonl
rm -f $PPS_MARGINS_FILE
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1
tail -4 $O/gpu_suite.log | cut -c1-300
unset PPS_MARGINS_FILE
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
cut -c1-400 $O/bench.json
timeout 200 python tools/fused_check.py --converge 256 > $O/fused_converge.log 2>&1; cut -c1-600 $O/fused_converge.log
