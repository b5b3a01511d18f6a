"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (the reference's CPU loop
through oracle/_ref or the C restatement) prints one JSON line with the keys the driver reads; the GPU arm refuses to
run without a CUDA device instead of falling back to anything."""
import json
import os
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


def test_clock_sampler_windows_by_arrival_time(tmp_path, monkeypatch):
    """The clocks block of the bench line: nvidia-smi is started with the device and its lines are stamped on arrival, so a
    timed region shorter than nvidia-smi's start-up (the 8-GPU strong-scaling run: 0.3 s) still gets the samples that
    fall inside it — or, failing that, the nearest one with its distance.  A stand-in nvidia-smi plays the tool here."""
    import time

    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.5\ni=0\nwhile true; do echo \"0, 1965, 1980, 600.$i, Not Active, Not Active, "
                    "Not Active, Active\"; i=$((i+1)); sleep 0.1; done\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    sys.path.insert(0, str(ROOT))
    import bench

    s = bench.ClockSampler(0)
    time.sleep(0.9)
    with s.window():
        time.sleep(0.45)
    got = s.summary()
    s.close()
    assert 2 <= got["samples"] <= 12 and got["sm_mhz"] == 1965.0 and got["sm_max_mhz"] == 1980.0
    assert got["reasons"] == ["sw_power_cap"] and "nearest_sample_s" not in got

    s = bench.ClockSampler(0)       # the window closes before the first line arrives
    with s.window():
        time.sleep(0.02)
    got = s.summary()
    s.close()
    assert got["samples"] == 1 and 0.2 < got["nearest_sample_s"] < 1.5

    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    s = bench.ClockSampler(0)       # no nvidia-smi at all: an empty block, no exception, no waiting
    with s.window():
        pass
    assert s.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    s.close()
