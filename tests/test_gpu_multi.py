"""Multi-GPU parity on hardware (needs >= 2 GPUs on the box; skipped otherwise): torchrun, NCCL, one process per
GPU -- see tests/mgpu_worker.py for what is checked.  Run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi`."""

import os
import subprocess
import sys

import pytest

from aurora_b200 import _native as N

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world: int):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and f"MGPU_OK {world}" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_two_gpu_sharded_search_matches_oracle():
    n = N.load().aur_device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, this box has {n}")
    _run(2)


def test_all_gpus_sharded_search_matches_oracle():
    n = N.load().aur_device_count()
    if n < 4:
        pytest.skip(f"needs >= 4 GPUs, this box has {n}")
    _run(min(n, 8))
