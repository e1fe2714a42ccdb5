O=gpurun_out/c17
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1
tail -6 $O/gpu_suite.log | cut -c1-400
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log | cut -c1-200
