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


@pytest.mark.parametrize("scheme_name", ["ab", "aa"])
@pytest.mark.parametrize("shape", [(64, 8, 8), (130, 8, 6), (516, 4, 5)])
def test_vec2_step_matches_the_64bit_step(shape, scheme_name):
    # D3Q19_VEC2=1: k_step_ab2 / k_step_aa2 (two nodes per thread, 128-bit accesses) run the same collide_fast on the same values;
    # on the host build it is bit-identical, on the device nvcc may contract the inlined arithmetic differently
    import numpy as np
    from oracle import oracle as orc
    pkg = entry.load_package()
    capi = pkg.capi
    nx, ny, nz = shape
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
    sims = []
    for v2 in (False, True):
        os.environ.pop("D3Q19_VEC2", None)
        if v2:
            os.environ["D3Q19_VEC2"] = "1"
        try:
            sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=capi.SCHEME_AB if scheme_name == "ab" else capi.SCHEME_AA,
                                  math_mode=capi.MATH_FAST)
        finally:
            os.environ.pop("D3Q19_VEC2", None)
        sim.FORCING()
        sim.upload_f(w.get_f())
        sims.append(sim)
    a, b = sims
    oa, ob = np.empty((nz, ny, nx, 19)), np.empty((nz, ny, nx, 19))
    w.macrovar(); w.collision_MRT()
    a.run_device(1); b.run_device(1)
    a.download_f(oa); b.download_f(ob)
    scale = np.max(np.abs(w.get_f()))
    assert np.max(np.abs(ob - w.get_f())) < 1e-12 * scale          # BASELINE.json: 1 step
    assert np.max(np.abs(ob - oa)) < 1e-14 * scale
    a.run_device(9); b.run_device(9)
    a.download_f(oa); b.download_f(ob)
    assert np.max(np.abs(ob - oa)) < 1e-13 * np.max(np.abs(oa))
    a.close(); b.close(); w.close()
