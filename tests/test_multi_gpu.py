"""N>1 on real GPUs: launches tests/multi_gpu_worker.py with one process per GPU. Skipped on a box with one GPU
(the 2-rank host logic is covered on CPU by tests/test_host_logic.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("allreduce", ["peer_memory", "nccl"])
def test_sharded_path_on_all_gpus(allreduce):
    """peer_memory: the all-reduce fused into the Gram kernels over NVLink peer stores (default);
    nccl: ncclAllReduce + copy + synchronise (what runs when the exchange buffers are not mapped)"""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("one GPU on this box")
    nproc = 2 if ngpu < 4 else (4 if ngpu < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    env = dict(os.environ)
    if allreduce == "nccl":
        env["ITSOLV_P2P_ALLREDUCE"] = "-1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(nproc):
        assert f"rank {k}/{nproc} ok" in r.stdout
