# call 15 (1 GPU): the extended drop-in driver tests + ncu refresh of the final build
O=gpurun_out/c15
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -q > $O/driver_suite.log 2>&1
tail -8 $O/driver_suite.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file $O/launches_bench512.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-gate > $O/bench_under_ncu.log 2>&1
for k in PreSUpdate PrePUpdate; do
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 10 -c 1 -f -o $O/prof_r02_512_final_$k python tools/probe.py iters 512 30 > $O/ncu_$k.log 2>&1
done
ls -la $O
