O=gpurun_out/c22
mkdir -p $O
export PPS_MARGINS_FILE=$O/margins.jsonl
# new / affected tests first (alpaka fixtures, global nested BiCGSTAB, drop-in driver), then the rest of the suite
timeout 150 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_next.py tests/test_gpu_driver.py tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $O/gpu_suite.log 2>&1
tail -25 $O/gpu_suite.log | cut -c1-400
