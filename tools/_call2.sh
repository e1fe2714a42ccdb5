# 2-GPU: the transport / overlap variants that have not had a parity run since the peer transport was generalised
O=gpurun_out/c20
mkdir -p $O
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
i=0
for v in "PPS_OVERLAP=2 PPS_HALO_P2P=0" "PPS_OVERLAP=3" "PPS_OVERLAP=3 PPS_ALLREDUCE_P2P=1" "PPS_HALO_P2P=0 PPS_ALLREDUCE_P2P=1"; do
  for f in fusecmp cheb; do
    i=$((i+1))
    env $v timeout 100 $TORCHRUN --nproc-per-node 2 tools/mg_check.py 1 1 2 $f > $O/mg_$i.json 2> $O/mg_$i.err
    echo "$v $f rc=$? $(cut -c1-260 $O/mg_$i.json)"
  done
done
