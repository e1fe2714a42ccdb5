O=gpurun_out/c21
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1
tail -3 $O/gpu_suite.log | cut -c1-300
