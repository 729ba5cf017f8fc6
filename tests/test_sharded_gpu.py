"""Sharded DLRM (table-wise + row-wise model parallel) == single-GPU DLRM; needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["peer/owner", "peer/direct", "p2p", "nccl"])
def test_sharded_dlrm_matches_single_gpu(rtf, mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (bench.py runs the same check in-process at every N > 1)")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29711",
           os.path.join(ROOT, "tests", "mgpu_check.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                         env=dict(os.environ, RTF_EXCHANGE=mode.split("/")[0],
                                  RTF_PEER_GATHER=mode.split("/")[-1]))
    assert res.returncode == 0 and "mgpu_check ok" in res.stdout, res.stdout[-3000:]
