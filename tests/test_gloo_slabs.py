"""N>1 host path on CPU: two processes, gloo backend (see tests/gloo_worker.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_process_slab_exchange_over_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "gloo_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "GLOO_SLABS_OK" in res.stdout, res.stdout[-3000:]


def test_reference_arm_under_torchrun_prints_one_line():
    # the driver launches `bench.py --impl reference` like the product arm (torchrun, N ranks): rank 0 alone runs the reference's
    # CPU path and prints the one JSON line, the other ranks exit 0 without work; no GPU is touched
    import json
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "3", "--workload", "32x8x8"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-3000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 3
    assert d["metric"] == "MLUPS (fp64)" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) == {"workload", "per_gpu", "precision", "particles"}          # the object both arms share
