"""GPU, 2+ devices: tests/multi_gpu_check.py -- every distributed mode (IterativeRecon fixed / radial line of sight,
MultigridRecon, randoms, TSC, PCS, the slab transform with all exchange variants, reconstruct_dist from an
interleaved split) against the oracle -- launched under torchrun from inside the suite, so that a box with several
GPUs checks the multi-rank path with `pytest -m gpu` alone.  Skipped on a single GPU, where tests/test_gpu_dist.py
runs the same code path with one rank.  Logs of the check at 2 and 8 ranks: profiles/r2_multi_gpu_check_*.log."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_multi_rank_parity_under_torchrun():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the multi-rank check needs two")
    ranks = 2 if n < 4 else (4 if n < 8 else 8)
    env = dict(os.environ, MGC_TSC="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "tests" / "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, tail
