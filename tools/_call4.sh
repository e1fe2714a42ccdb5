# 4-GPU call: parity + bench at N = 4 with the final build; reproducer attempt for the in-kernel allreduce
O=gpurun_out/c14
mkdir -p $O
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
export PPS_MARGINS_FILE=$PWD/$O/parity_margins_4gpu.jsonl
timeout 120 $TORCHRUN --nproc-per-node 4 tools/mg_check.py 1 1 4 fusecmp > $O/mg_114.json 2> $O/mg_114.err
timeout 120 $TORCHRUN --nproc-per-node 4 tools/mg_check.py 1 2 2 cheb > $O/mg_122.json 2> $O/mg_122.err
PPS_ALLREDUCE_P2P=2 timeout 120 $TORCHRUN --nproc-per-node 4 tools/mg_check.py 1 1 4 > $O/mg_114_arp2p.json 2> $O/mg_114_arp2p.err
cat $O/mg_114.json $O/mg_122.json $O/mg_114_arp2p.json | cut -c1-500
unset PPS_MARGINS_FILE
timeout 300 $TORCHRUN --nproc-per-node 4 bench.py --gpus 4 --steps 2 --warmup 3 --watchdog 280 > $O/bench4.json 2> $O/bench4.err
cut -c1-900 $O/bench4.json; grep -v "^\[W\|^W1017\|^\*\*\*\|^$" $O/bench4.err | tail -2 | cut -c1-300
PPS_ALLREDUCE_P2P=2 timeout 200 $TORCHRUN --nproc-per-node 4 bench.py --gpus 4 --steps 1 --warmup 1 --watchdog 180 --no-cpu-baseline > $O/bench4_arp2p.json 2> $O/bench4_arp2p.err
cut -c1-600 $O/bench4_arp2p.json; grep "PARITY" $O/bench4_arp2p.err | head -1 | cut -c1-500
