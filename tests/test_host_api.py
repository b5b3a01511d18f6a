"""Builds and runs the C++ host-API tests (tests/host/test_host_api.cpp): the reference's public classes
(SceneDescription, PerspectiveCamera, FrameTiling, FrameBuffer, RGB, RenderSession) mirrored in include/cornelis/."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "cornelis_b200" / "lib"


@pytest.fixture(scope="module")
def host_test_binary(tmp_path_factory):
    from cornelis_b200 import build
    build.build_all()
    exe = tmp_path_factory.mktemp("host") / "test_host_api"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "host" / "test_host_api.cpp"), "-o", str(exe), f"-L{LIB}", "-lcorneliscore",
                    "-lcornelis_cuda", f"-Wl,-rpath,{LIB}", "-pthread"], check=True)
    return exe


def test_host_api_cpu(host_test_binary):
    import torch
    mode = "cpu" if torch.cuda.is_available() else "cpu-nodevice"
    r = subprocess.run([str(host_test_binary), mode], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cli_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(LIB / "cornelis"), "--width", "32", "--height", "32", "--spp", "1", "--no-save"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_host_api_gpu(host_test_binary):
    import torch
    r = subprocess.run([str(host_test_binary), "gpu", str(torch.cuda.device_count())], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_renders_reference_default_scene(tmp_path):
    out = tmp_path / "cornell.png"
    r = subprocess.run([str(LIB / "cornelis"), "--spp", "16", "--output", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "512x512, 16 spp" in r.stdout and out.exists() and out.stat().st_size > 512 * 512 * 3


@pytest.mark.gpu
def test_cli_many_spheres_grid_and_exhaustive_agree(tmp_path):
    """The CLI's many-sphere scene (BASELINE configs[3] style) through the grid and through the exhaustive scan: the
    same paths, hence the same statistics line apart from the timings."""
    lines = []
    for accel in ("grid", "none"):
        r = subprocess.run([str(LIB / "cornelis"), "--scene", "spheres", "--spheres", "2000", "--width", "96", "--height",
                            "54", "--spp", "8", "--max-depth", "64", "--accel", accel, "--no-save"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        lines.append(r.stdout.strip().splitlines()[-1])
    tail = [ln.split("Mrays/s,")[1] for ln in lines]     # "x rays/sample, deepest path y"
    assert tail[0] == tail[1], lines
