#!/usr/bin/env bash
# One-call GPU validation plans for the next round.  Every step has its own timeout and writes its log under
# gpurun_out/r2/, so one gpurun call brings back everything even when a step fails or is cut off.
#
#   gpurun --timeout 1500 -- 'bash tools/r2_validate.sh one'        # 1 GPU, about 20 minutes
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_validate.sh two'   # 2 GPUs
#   gpurun --gpus 8 --timeout 420 -- 'bash tools/r2_validate.sh eight' # 8 GPUs: SHORT on purpose (charged 8x)
#
# Never put a multi-rank command under ncu, never combine the 8-GPU bench with PPS_OVERLAP=2 (deadlocked in round 1).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2
mkdir -p "$OUT"
step() {   # step <name> <timeout seconds> <command...>
    local name=$1 limit=$2
    shift 2
    echo "== $name (limit ${limit}s)" | tee -a "$OUT/summary.txt"
    local t0=$SECONDS
    timeout --kill-after=20 "$limit" "$@" >"$OUT/$name.log" 2>&1
    local rc=$?
    echo "   rc=$rc  $((SECONDS - t0))s  $(tail -n 1 "$OUT/$name.log" | cut -c1-200)" | tee -a "$OUT/summary.txt"
}
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"

case "${1:-one}" in
one)
    step smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
    # 1. the verified default suite must still be green (launchers now go through h->actl)
    step gpu_suite 900 python -m pytest tests -m gpu -x -q -k "not multi and not driver"
    # 2. what round 1 wrote without a GPU: orderNeumanBcs = 1, Chebyshev as main solver, nested Krylov, batched ghosts
    PPS_TEST_EXPERIMENTAL=1 step next_paths 600 python -m pytest tests/test_gpu_next.py -m gpu -q
    PPS_TEST_EXPERIMENTAL=1 step batched_ghosts 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "batched or graph"
    # launch-bound shipped default problem (128x128x256, Chebyshev): stream launches vs graph replay vs graph + batched ghosts
    step default_stream 120 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1
    PPS_GRAPH=1 step default_graph 120 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1
    PPS_GRAPH=1 PPS_BATCH_GHOSTS=1 step default_graph_batched 120 parallelpoissonsolver_b200/driver/solverPoisson 1 1 1
    step drivers 600 python -m pytest tests/test_gpu_driver.py -m gpu -q
    # 3. the 17-pass schedule: reproduce / bisect the 512^3 repeat-solve anomaly
    step fused_fixed 300 python tools/fused_check.py 128 256 512
    step fused_converge 500 python tools/fused_check.py --converge 128 256 512
    PPS_FUSE_P=0 step fused_converge_noP 300 python tools/fused_check.py --converge 512
    PPS_FUSE_S=0 step fused_converge_noS 300 python tools/fused_check.py --converge 512
    step fused_racecheck 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/fused_check.py 64
    step fused_memcheck 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/fused_check.py 64
    # 4. bench lines
    step bench_default 400 python bench.py --steps 3 --warmup 3
    PPS_FUSE=2 step bench_fused 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
    ;;
two)
    step multi_suite 400 python -m pytest tests/test_gpu_multi.py -m gpu -q
    PPS_TEST_EXPERIMENTAL=1 step transports 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -k transport
    step overlap_default 200 $TORCHRUN --nproc-per-node 2 tools/overlap_probe.py 1024 1024 256 150
    step overlap_p2p 200 $TORCHRUN --nproc-per-node 2 tools/overlap_probe.py 1024 1024 256 150 --p2p
    step bench2 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350
    PPS_HALO_P2P=1 PPS_OVERLAP=3 step bench2_p2p_inkernel 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 --no-cpu-baseline
    PPS_HALO_P2P=1 PPS_ALLREDUCE_P2P=1 step bench2_p2p 400 $TORCHRUN --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 --watchdog 350 --no-cpu-baseline
    ;;
eight)
    step bench8 200 $TORCHRUN --nproc-per-node 8 bench.py --gpus 8 --steps 2 --warmup 3 --watchdog 180 --no-cpu-baseline
    step precond8 150 $TORCHRUN --nproc-per-node 8 tools/precond_compare.py 768
    ;;
*)
    echo "usage: $0 one|two|eight" >&2
    exit 2
    ;;
esac
cat "$OUT/summary.txt"
