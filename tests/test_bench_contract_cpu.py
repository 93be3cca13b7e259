"""Driver contract of bench.py, the parts that run without a GPU: the reference arm prints ONE JSON line with the agreed
keys (tiny k so it takes seconds), the product arm refuses to run without CUDA instead of falling back, and the
schedule is the 38 MSM + 59 NTT of SURVEY.md App. C."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--k", "10"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is False and line["unit"] == "s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the product arm runs; that is the -m gpu suite's business
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--k", "10"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_schedule_shape():
    sys.path.insert(0, ROOT)
    import bench

    units = bench.schedule()
    kinds = [u[1] for u in units]
    assert kinds.count("msm") == 38 and kinds.count("intt") == 29 and kinds.count("coset") == 29 and kinds.count("ext_intt") == 1
    assert bench.algorithmic_bytes(22) == 10400 * (1 << 22) + 38 * 96
