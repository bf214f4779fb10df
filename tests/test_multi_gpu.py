"""Multi-GPU parity (gpu marker): runs tests/mgpu_check.py under torchrun on 2 (and, when the box has
them, 4 and 8) GPUs.  Skipped on single-GPU boxes; the host-side logic is covered on CPU by
tests/test_decomp_gloo.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_decomposed_matches_single_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT,
                       env=dict(os.environ, MGPU_NCELL="8", MGPU_CUT="8.0"))
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all cases passed" in r.stdout


@pytest.mark.parametrize("world", [2, 8])
def test_decomposed_md_step_matches_single_gpu(world):
    """rigid integrator + pair style + KSpace, all decomposed, against the single-GPU trajectory"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29650 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "mgpu_md_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all md cases passed" in r.stdout
