# 2-GPU call: NCCL-path parity incl. fused-vs-split bitwise, transport variants, overlap probes, bench lines
O=gpurun_out/c7
mkdir -p $O
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
export PPS_MARGINS_FILE=$PWD/$O/parity_margins_2gpu.jsonl
rm -f $PPS_MARGINS_FILE
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/multi_suite.log 2>&1
tail -3 $O/multi_suite.log
PPS_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k transport > $O/transports.log 2>&1
tail -15 $O/transports.log | cut -c1-300
unset PPS_MARGINS_FILE
timeout 200 $TORCHRUN --nproc-per-node 2 tools/overlap_probe.py 1024 1024 256 150 --p2p > $O/overlap_fused.json 2> $O/overlap_fused.err
PPS_FUSE=1 timeout 200 $TORCHRUN --nproc-per-node 2 tools/overlap_probe.py 1024 1024 256 150 --p2p > $O/overlap_split.json 2> $O/overlap_split.err
cat $O/overlap_fused.json $O/overlap_split.json | cut -c1-1500
timeout 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 > $O/bench2.json 2> $O/bench2.err
PPS_ALLREDUCE_P2P=1 timeout 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 --no-cpu-baseline > $O/bench2_arp2p.json 2> $O/bench2_arp2p.err
PPS_FUSE=1 timeout 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 --no-cpu-baseline > $O/bench2_split.json 2> $O/bench2_split.err
for f in bench2 bench2_arp2p bench2_split; do cut -c1-700 $O/$f.json; tail -2 $O/$f.err | cut -c1-300; done
