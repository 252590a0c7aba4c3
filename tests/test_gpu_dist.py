"""GPU, >= 2 devices: run tests/dist_gpu_check.py under torchrun (NCCL).  Skipped on a single-GPU box; the host-side
logic is covered on CPU by tests/test_dist_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("n,reach", [(2, "auto"), (2, "1"), (4, "1"), (8, "1")])
def test_multi_gpu_parity(n, reach):
    """DimShard, RowPartition (NCCL block exchange) and PeerRowPartition (exchange fused into the kernels over NVLink peer
    memory) on n ranks: fused steps equal the reference fixtures, row partitions are bit-identical to one GPU.
    reach = "1": the engines also restrict layer L-1 / backward hops 1-2 to the batch's one-hop mask (B200REC_REACH, which
    the C4 bench turns on by itself) -- same fixtures, same bars."""
    if torch.cuda.device_count() < n:
        pytest.skip("needs >= %d GPUs" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(29611 + n), os.path.join(HERE, "dist_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, B200REC_REACH=reach))
    assert out.returncode == 0 and "DIST_GPU_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
