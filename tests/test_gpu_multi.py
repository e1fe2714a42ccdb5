"""Multi-GPU parity (one process per GPU, NCCL face exchange + allreduce) against the oracle on the same rank
layout.  Needs >= 2 GPUs; skipped on a single-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(layout, flags=(), env=None):
    n = layout[0] * layout[1] * layout[2]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "mg_check.py"), *map(str, layout), *flags]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(lines[-1])
    assert out["ok"], out
    return out


@pytest.mark.parametrize("layout", [(1, 1, 2), (2, 1, 1), (1, 2, 1)])
@pytest.mark.parametrize("flags", [("fusecmp",), ("cheb",), ("cg",)])
def test_two_gpus(layout, flags):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(layout, flags)


@pytest.mark.parametrize("layout", [(1, 1, 4), (2, 2, 1), (1, 2, 2)])
def test_four_gpus(layout):
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run(layout, ("cheb",))


@pytest.mark.parametrize("layout", [(1, 1, 8), (2, 2, 2)])
def test_eight_gpus(layout):
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run(layout, ("fusecmp",))


# transports and overlap schedules (validated on 2 B200s in round 2).  Default: peer-memory face pushes (CUDA IPC + copy engines,
# PPS_HALO_P2P=1), three-stream overlap (PPS_OVERLAP=1), NCCL allreduce.
@pytest.mark.parametrize("env", [{"PPS_HALO_P2P": "0"}, {"PPS_ALLREDUCE_P2P": "1"}, {"PPS_HALO_P2P": "0", "PPS_ALLREDUCE_P2P": "1"},
                                 {"PPS_OVERLAP": "0"}, {"PPS_OVERLAP": "2", "PPS_HALO_P2P": "0"}, {"PPS_OVERLAP": "3"},
                                 {"PPS_OVERLAP": "3", "PPS_ALLREDUCE_P2P": "1"}])
@pytest.mark.parametrize("flags", [("fusecmp",), ("cheb",)])
def test_two_gpus_slab_transport_variants(env, flags):
    """the same parity check over NCCL send/recv instead of the peer-memory pushes, with the in-kernel allreduce over peer
    memory, with the serial exchange, and with the in-kernel-wait overlaps (PPS_OVERLAP=2 over NCCL, =3 over the SM-free
    peer transport)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run((1, 1, 2), flags, env)
