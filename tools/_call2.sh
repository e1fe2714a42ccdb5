# 2-GPU call: NCCL / peer-transport parity (fused vs split bitwise), overlap probe, bench lines
O=gpurun_out/c10
mkdir -p $O
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
export PPS_MARGINS_FILE=$PWD/$O/parity_margins_2gpu.jsonl
rm -f $PPS_MARGINS_FILE
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "test_two_gpus and not transport" > $O/multi_suite.log 2>&1
tail -3 $O/multi_suite.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "transport and fusecmp" > $O/transports.log 2>&1
tail -3 $O/transports.log | cut -c1-300
unset PPS_MARGINS_FILE
timeout 300 $TORCHRUN --nproc-per-node 2 tools/overlap_probe.py 1024 1024 256 150 --p2p > $O/overlap.json 2> $O/overlap.err
cut -c1-1500 $O/overlap.json
timeout 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 > $O/bench2.json 2> $O/bench2.err
PPS_ALLREDUCE_P2P=1 timeout 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 --no-cpu-baseline > $O/bench2_arp2p.json 2> $O/bench2_arp2p.err
for f in bench2 bench2_arp2p; do cut -c1-900 $O/$f.json; grep -v "^\[W\|^W1017\|^\*\*\*\|^$" $O/$f.err | tail -2 | cut -c1-300; done
