"""The C-ABI library ITSELF on the GPU-less build box: tests/host/make_hostsim.py compiles csrc/d3q19_api.cu and the
kernels with g++ against a synchronous stand-in for the CUDA runtime (tests/host/fake/, every kernel launch rewritten
into a loop nest / fibers, NCCL = mailboxes between threads) and the GPU test files are run against it in
subprocesses, unchanged, through the same ctypes binding and the same C++ driver:

* single rank: tests/test_gpu_parity.py, test_golden.py, test_gpu_particles.py, test_zz_cpp_driver.py -- handle life
  cycle, transfers in every storage phase, the shim state machine and its download policy, pre-relaxation, reductions;
* 2 and 3 ranks (threads): tests/host/hostsim_mrank_worker.py -- slab geometry, the face exchange and the send-back
  after odd in-place steps, peer-memory connect and flag protocol (fused, split, copy-engine put),
  all-reduced scalars, the particle link partition, force all-reduce and the refill source exchange; all bit for
  bit against the single-domain oracle.

What this cannot show is anything about time or about ordering between CUDA streams (tests/test_halo_schedule_model.py
covers the latter).  TEST INFRASTRUCTURE: the package never builds, finds or loads the host-sim library; the GPU box
runs the same files against libd3q19b200.so.
"""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostsim():
    spec = importlib.util.spec_from_file_location("make_hostsim", os.path.join(ROOT, "tests", "host", "make_hostsim.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return {"lib": m.build(), "driver": m.build_driver()}


def run(cmd, hostsim, timeout=900, **extra):
    env = dict(os.environ, D3Q19_LIB=hostsim["lib"], D3Q19_DRIVER=hostsim["driver"], **extra)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_single_rank_gpu_test_files_pass_on_the_host_sim(hostsim):
    # left out: the 1000-step cases (minutes on a CPU); the particle force / moving cases (the multi-rank worker below
    # runs the same sequence through the same entry points, tests/test_kernels_host.py the kernels)
    res = run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_golden.py", "tests/test_gpu_particles.py",
               "tests/test_zz_cpp_driver.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
               "-k", "not 1000_steps and not moving_particles and not ibb_and_force and not momentum_exchange and not faxen and not many_particles"], hostsim)
    tail = res.stdout[-3000:]
    assert res.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail and "error" not in tail.lower(), tail


@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_orchestration_on_the_host_sim(hostsim, world):
    # 3 ranks: every case of the worker; 2 ranks (both neighbours are the same rank): one uneven case per transport
    extra = {"HOSTSIM_SHORT": "1"} if world == 2 else {}
    sections = ["fluid", "shim", "particles", "random", "benchparity"] if world == 2 else []            # none named = all of them
    res = run([sys.executable, os.path.join("tests", "host", "hostsim_mrank_worker.py"), str(world)] + sections, hostsim, **extra)
    assert res.returncode == 0 and "HOSTSIM_MRANK_OK" in res.stdout, res.stdout[-4000:]


def test_smoke_entry_point_on_the_host_sim(hostsim):
    # __graft_entry__.smoke() is what the driver runs on the B200 before the bench: its own logic must not be what fails
    res = run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], hostsim)
    assert res.returncode == 0 and "smoke ok" in res.stdout, res.stdout[-3000:]


@pytest.mark.parametrize("extra", [[], ["--scheme", "aa", "--device-init"], ["--workload", "40x24x24", "--particles", "2", "--rad", "4"]])
def test_bench_product_arm_prints_its_contract_line(hostsim, extra):
    # bench.py's own logic (the one JSON line the driver reads) against the host-sim; numbers are meaningless here
    import json
    args = ["--workload", "32x8x8", "--steps", "6", "--warmup", "3", "--cpu-steps", "1"] + extra
    res = run([sys.executable, os.path.join("tests", "host", "bench_on_hostsim.py")] + args, hostsim)
    assert res.returncode == 0, res.stdout[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "impl"):
        assert k in d, k
    assert d["impl"] == "ours" and d["metric"] == "MLUPS (fp64)" and d["dtype"] == "f64" and d["n_gpus"] == 1
    assert d["steps"] == 6 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] >= 6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["bytes_per_node"] == 304.0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and "peak_source" in r
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    if "--particles" in extra or "--device-init" in extra:
        assert d["e2e"] is None                      # those runs have no host-buffer leg
    else:
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] == 19 * 8 * 32 * 8 * 8 / 6 and e["d2h_bytes_per_step"] > 32
    assert d["config"]["workload"].startswith("D3Q19 MRT channel")
    pc = d["parity_check"]                            # the reference's golden vector, checked before the timed region
    assert pc["bit_exact"] is True and pc["ranks"] == 1 and [t["scheme"] for t in pc["schemes"]] == ["aa", "ab"]


def test_full_size_property_tests_on_the_host_sim(hostsim):
    # the logic of tests/test_zzz_gpu_fullsize.py on a shrunk channel
    res = run([sys.executable, "-m", "pytest", "tests/test_zzz_gpu_fullsize.py", "-m", "gpu",
               "-q", "-p", "no:cacheprovider"],
              hostsim, D3Q19_TEST_FULLSIZE="64x16x16")
    tail = res.stdout[-3000:]
    assert res.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail


def test_random_call_sequences_on_the_host_sim(hostsim):
    # random but legal orders of the raw C-ABI calls and of the driver-facing shim calls, mirrored on the oracle
    res = run([sys.executable, os.path.join("tests", "host", "hostsim_random_calls.py"), "5"], hostsim)
    assert res.returncode == 0 and "HOSTSIM_RANDOM_OK" in res.stdout, res.stdout[-4000:]
