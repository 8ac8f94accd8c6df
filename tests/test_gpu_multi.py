"""z-slab decomposition over several GPUs: bit-identical to the single-domain oracle."""
import os
import subprocess
import sys

import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return entry.load_package().capi.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0 and "MGPU_PARITY_OK" in res.stdout, res.stdout[-4000:]
