"""N > 1 on real GPUs (skipped on a 1-GPU box): sharded training with the NCCL count-slab all-reduce and
sharded scoring without a collective give the single-GPU results.  Worker: tests/mgpu_worker.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_gpus_match_one():
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(HERE, "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
