"""Real multi-GPU slab run (one process per GPU, NCCL send/recv + allreduce) against the single-domain engine.
Skipped on boxes with fewer than 2 GPUs; run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_slabs_match_single_domain(world, transport):
    """transport = peer: mailboxes in cudaIpc-mapped peer memory, step replayed as a CUDA graph (the default);
    nccl: ncclSend/ncclRecv + ncclAllReduce with eager launches (MDB200_NO_PEER=1)"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ)
    if transport == "nccl":
        env["MDB200_NO_PEER"] = "1"
    env["MDB200_EXPECT_TRANSPORT"] = {"peer": "3", "nccl": "2"}[transport]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world + (10 if transport == "nccl" else 0)), os.path.join(ROOT, "tests", "mp_slab_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, cwd=ROOT, env=env)
    out = p.stdout.decode()
    assert p.returncode == 0 and "MULTIRANK OK" in out, out[-4000:]
