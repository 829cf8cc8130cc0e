"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on the 1-GPU box): the NCCL leg exchange of the
fused plan, checked by tools/multi_gpu_check.py against the oracle (small sizes) and against a
single-rank plan (nside 1024), both spins, both directions."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("host_blocks", ["1", "3"])
@pytest.mark.parametrize("nranks", [2, 3, 4, 6, 8])
def test_nccl_ranks_match_single_rank_and_oracle(nranks, host_blocks):
    """host_blocks 1: one exchange per transform; 3: ring blocks pipelined (Legendre | exchange | ring FFT |
    host copy), which the host-array API uses at large sizes."""
    import torch

    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + nranks),
           os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, HPSHT_HOST_BLOCKS=host_blocks)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
