mkdir -p gpurun_out/c5
export PPS_MARGINS_FILE=$PWD/gpurun_out/c5/parity_margins.jsonl
rm -f $PPS_MARGINS_FILE
timeout 900 python tools/fused_debug.py 512 > gpurun_out/c5/fused_debug.jsonl 2> gpurun_out/c5/fused_debug.err
timeout 600 python tools/fused_check.py --converge 256 512 > gpurun_out/c5/fused_converge.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "not multi" > gpurun_out/c5/gpu_suite.log 2>&1
tail -3 gpurun_out/c5/gpu_suite.log
PPS_FUSE=2 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c5/bench_fused.json 2> gpurun_out/c5/bench_fused.err
cut -c1-300 gpurun_out/c5/fused_debug.jsonl
cut -c1-1500 gpurun_out/c5/fused_converge.log
cut -c1-600 gpurun_out/c5/bench_fused.json
