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


@pytest.fixture(scope="module")
def cli():
    """The CLI, built in-tree by build.py (a build artefact, not tracked)."""
    from cornelis_b200 import build
    build.build_all()
    exe = LIB / "cornelis"
    assert exe.exists()
    return exe


def test_host_api_cpu(host_test_binary):
    import torch
    mode = "cpu" if torch.cuda.is_available() else "cpu-nodevice"
    r = subprocess.run([str(host_test_binary), mode], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def _decode_png_rgb8(path):
    """A PNG decoder for 8-bit RGB, non-interlaced files: chunks, zlib, the five row filters."""
    import struct
    import zlib

    import numpy as np
    data = Path(path).read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, width, height = 8, b"", 0, 0
    while pos < len(data):
        (n,), tag = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0], tag
        if tag == b"IHDR":
            width, height, depth, colour, _, _, interlace = struct.unpack(">IIBBBBB", body)
            assert (depth, colour, interlace) == (8, 2, 0)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(height, 1 + 3 * width)
    out = np.zeros((height, 3 * width), np.int32)
    for j in range(height):
        kind, row = int(raw[j, 0]), raw[j, 1:].astype(np.int32)
        up = out[j - 1] if j else np.zeros(3 * width, np.int32)
        for i in range(3 * width):
            a = out[j, i - 3] if i >= 3 else 0
            b, c = up[i], (up[i - 3] if i >= 3 else 0)
            p = a + b - c
            pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
            pred = [0, a, b, (a + b) // 2, a if (pa <= pb and pa <= pc) else b if pb <= pc else c][kind]
            out[j, i] = (row[i] + pred) & 255
    return out.reshape(height, width, 3).astype(np.uint8), len(data)


def test_save_image_writes_a_real_png(host_test_binary, tmp_path, port_oracle):
    """saveImage (Render.cpp:257-265): toSRGB, quantizeTo8bit, PNG.  The file must decode to exactly the quantised
    display values (oracle's toSRGB restatement) and be compressed (the rows are filtered and deflated)."""
    import numpy as np
    W, H = 97, 41
    out = tmp_path / "gradient.png"
    r = subprocess.run([str(host_test_binary), "png", str(out), str(W), str(H)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    image, size = _decode_png_rgb8(out)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    linear = np.stack([i / np.float32(W), j / np.float32(H), ((i * 7 + j * 13) % 32) / np.float32(16)], axis=2)
    expect = port_oracle.to_srgb8(linear.reshape(-1, 3).astype(np.float32)).reshape(H, W, 3)
    assert np.array_equal(image, expect)
    assert size < W * H * 3  # smaller than the raw pixels: the two smooth channels compress


def test_cli_fails_loudly_without_gpu(cli):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(cli), "--width", "32", "--height", "32", "--spp", "1", "--no-save"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_host_api_gpu(host_test_binary):
    import torch
    r = subprocess.run([str(host_test_binary), "gpu", str(torch.cuda.device_count())], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_renders_reference_default_scene(cli, tmp_path):
    out = tmp_path / "cornell.png"
    r = subprocess.run([str(cli), "--spp", "16", "--output", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "512x512, 16 spp" in r.stdout and out.exists()
    image, _ = _decode_png_rgb8(out)
    assert image.shape == (512, 512, 3) and image[:256].mean() > image[256:].mean() * 0.5 and image.max() == 255


@pytest.mark.gpu
def test_cli_many_spheres_grid_and_exhaustive_agree(cli, tmp_path):
    """The CLI's many-sphere scene (BASELINE configs[3] style) through the grid and through the exhaustive scan: the
    same paths, hence the same statistics line apart from the timings."""
    lines = []
    for accel in ("grid", "none"):
        r = subprocess.run([str(cli), "--scene", "spheres", "--spheres", "2000", "--width", "96", "--height",
                            "54", "--spp", "8", "--max-depth", "64", "--accel", accel, "--no-save"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        lines.append(r.stdout.strip().splitlines()[-1])
    tail = [ln.split("Mrays/s,")[1] for ln in lines]     # "x rays/sample, deepest path y"
    assert tail[0] == tail[1], lines
