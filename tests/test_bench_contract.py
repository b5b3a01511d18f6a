"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (the reference's CPU loop
through oracle/_ref or the C restatement) prints one JSON line with the keys the driver reads; the GPU arm refuses to
run without a CUDA device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("workload", ["cornell", "spheres"])
def test_reference_arm_json_line(workload, port_oracle, ref_oracle):  # the fixtures build the oracle libraries
    r = _run("--impl", "reference", "--workload", workload, "--width", "80", "--height", "40", "--steps", "1",
             "--warmup", "0")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "pixel-samples/s" and line["unit"] == "Msamples/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["gpu_launches"] == 0
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert "spp" in base["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CPU path" in (r.stdout + r.stderr)
