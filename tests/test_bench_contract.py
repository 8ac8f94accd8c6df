"""bench.py's reference arm (`--impl reference`) runs on the host cores alone, so its side of the driver's
contract can be checked here: one JSON line with the keys the driver reads, the same metric / unit / workload
naming as the GPU arm, and silent ranks > 0 under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = run(["--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "MLUPS (fp64)" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("D3Q19 MRT channel 64x32x32")


def test_reference_arm_other_ranks_are_silent():
    res = run(["--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1", "--warmup", "1"],
              env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    # no CPU fallback: on this GPU-less box the product arm must refuse, not compute
    import torch
    if torch.cuda.is_available():
        return
    res = run(["--workload", "c1", "--steps", "1", "--warmup", "1"])
    assert res.returncode != 0
    assert "CUDA device" in (res.stderr + res.stdout)
