"""Opt-in paths that have not become defaults yet; kept in a file that sorts last so that a failure here cannot mask
the established GPU tests.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_boundary_stream_is_bit_identical(world):
    # D3Q19_BOUNDARY_STREAM=1: the boundary launch of the NCCL transport next to the interior launch (d3q19_api.cu step_impl)
    if entry.load_package().capi.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MGPU_ONLY="bstream")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "MGPU_PARITY_OK" in res.stdout, res.stdout[-4000:]


@pytest.mark.parametrize("world", [2, 4])
def test_refill_is_decomposition_invariant(world):
    # d3q19_beads_filling exchanges the neighbours' boundary planes (19 populations) before the refill
    if entry.load_package().capi.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MGPU_ONLY="refill")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "MGPU_PARITY_OK" in res.stdout, res.stdout[-4000:]
