# 8-GPU call (charged 8x: keep it short).  Every step has its own timeout; bench has its own watchdog.
O=gpurun_out/c11
mkdir -p $O
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
export PPS_MARGINS_FILE=$PWD/$O/parity_margins_8gpu.jsonl
rm -f $PPS_MARGINS_FILE
timeout 150 $TORCHRUN --nproc-per-node 8 tools/mg_check.py 1 1 8 fusecmp > $O/mg_118.json 2> $O/mg_118.err
timeout 150 $TORCHRUN --nproc-per-node 8 tools/mg_check.py 2 2 2 fusecmp > $O/mg_222.json 2> $O/mg_222.err
cat $O/mg_118.json $O/mg_222.json | cut -c1-600
unset PPS_MARGINS_FILE
timeout 240 $TORCHRUN --nproc-per-node 8 bench.py --gpus 8 --steps 2 --warmup 3 --watchdog 200 > $O/bench8.json 2> $O/bench8.err
cut -c1-1200 $O/bench8.json; grep -v "^\[W\|^W1017\|^\*\*\*\|^$" $O/bench8.err | tail -2 | cut -c1-300
PPS_ALLREDUCE_P2P=1 timeout 200 $TORCHRUN --nproc-per-node 8 bench.py --gpus 8 --steps 2 --warmup 3 --watchdog 180 --no-cpu-baseline > $O/bench8_arp2p.json 2> $O/bench8_arp2p.err
cut -c1-700 $O/bench8_arp2p.json; grep -v "^\[W\|^W1017\|^\*\*\*\|^$" $O/bench8_arp2p.err | tail -2 | cut -c1-300
timeout 240 $TORCHRUN --nproc-per-node 8 tools/precond_compare.py 768 > $O/precond768.json 2> $O/precond768.err
cut -c1-2500 $O/precond768.json; grep "^#" $O/precond768.err | cut -c1-300
timeout 150 $TORCHRUN --nproc-per-node 8 tools/bandwidth_sweep.py 256 512 1024 1280 > $O/weak_sweep.jsonl 2> $O/weak_sweep.err
cut -c1-300 $O/weak_sweep.jsonl
timeout 150 $TORCHRUN --nproc-per-node 8 tools/overlap_probe.py 1024 1024 1024 100 > $O/overlap8.json 2> $O/overlap8.err
cut -c1-900 $O/overlap8.json
