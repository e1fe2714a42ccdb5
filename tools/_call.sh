O=gpurun_out/c23
mkdir -p $O
timeout 26 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_next.py -m gpu -q -p no:cacheprovider > $O/gpu_cheb_next.log 2>&1
tail -4 $O/gpu_cheb_next.log | cut -c1-300
timeout 14 python tools/nested_probe.py > $O/nested_probe.jsonl 2>$O/nested_probe.err
cat $O/nested_probe.jsonl; tail -2 $O/nested_probe.err
