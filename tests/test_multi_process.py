"""N>1 path: world_size-2 gloo test on the CPU (partition bookkeeping + exchange through a real process boundary)
and, when >= 2 GPUs are visible, the NCCL halo exchange inside the CUDA library."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(mode, nproc, timeout, extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "mp_worker.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.update(extra_env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_partitioned_oracle_gloo_world2():
    _torchrun("gloo-oracle", 2, 300)


def test_partitioned_type2_element_halo_gloo_world2():
    """face_integral_type 2 on a 2-way partition: element-data halo (parallel_data = element) through gloo, oracle only"""
    _torchrun("gloo-oracle-es2", 2, 300)


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "p2p-nograph", "p2p-unfused", "nccl", "nccl-inline"])
def test_nccl_halo_exchange_matches_serial(transport):
    """P-way result == serial result through the library's own halo exchange: remote stores into the neighbour's IPC-mapped
    receive buffer + flags from inside the one face launch of an evaluation, RK4 steps replayed as CUDA graphs (default);
    the same without graphs (PDES_GRAPH_MP=0); separate pack / signal / wait kernels on the communication stream
    (PDES_HALO_FUSED=0); ncclSend/ncclRecv (PDES_HALO_NCCL=1), and the latter with pack and shared-face flux on the
    compute stream (PDES_COMM_INLINE=1)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    env = {"p2p": {}, "p2p-nograph": {"PDES_GRAPH_MP": "0"}, "p2p-unfused": {"PDES_HALO_FUSED": "0"}, "nccl": {"PDES_HALO_NCCL": "1"}, "nccl-inline": {"PDES_HALO_NCCL": "1", "PDES_COMM_INLINE": "1"}}[transport]
    out = _torchrun("nccl-b200", 2 if n < 4 else 4, 600, env)
    assert "nccl-b200 ok" in out
