"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/cornelis_cuda.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "cornelis_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cornelis_cuda_\w+)\s*\(", text)))


def test_header_is_plain_c(tmp_path):
    """The boundary header must compile as C (no C++ types in the signatures)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "cornelis_cuda.h"\nint main(void){return (int)sizeof(cornelis_render_params) == 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), "-c", str(src),
                    "-o", str(tmp_path / "abi.o")], check=True)


def test_library_exports_every_declared_symbol():
    from cornelis_b200 import binding
    if not binding.LIB_PATH.exists():
        from cornelis_b200 import build
        build.build_cuda()
    lib = binding.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    assert sorted(binding.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cornelis_cuda_abi_version() == 2


def test_struct_layouts_match_header():
    from cornelis_b200 import binding
    assert C.sizeof(binding.CameraDesc) == 32 and C.sizeof(binding.MaterialDesc) == 44
    assert C.sizeof(binding.SphereDesc) == 20 and C.sizeof(binding.PlaneDesc) == 40
    assert C.sizeof(binding.RenderParams) == 48
    assert C.sizeof(binding.RenderStats) == 80


def test_no_device_fails_loudly():
    """Without a GPU the product must refuse to compute — there is no CPU path to fall back to."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from cornelis_b200 import binding, scenes
    with pytest.raises(binding.CornelisError) as e:
        binding.device_count()
    assert e.value.code == binding.ERR_NO_DEVICE
    with pytest.raises(binding.CornelisError) as e:
        binding.Scene(scenes.cornell_box())
    assert e.value.code == binding.ERR_NO_DEVICE and "no CPU path" in str(e.value)


def test_nccl_is_bound_at_run_time():
    """The exchange step's NCCL is loaded on first use (dlopen of libnccl.so.2), not linked: creating a communicator
    id needs no GPU, so this proves here that the library finds and binds NCCL.  The library itself must not list
    libnccl among its dependencies (a process that already carries torch's copy must not get a second one)."""
    import subprocess
    from cornelis_b200 import binding
    uid = binding.comm_unique_id()
    assert len(uid) == binding.COMM_ID_BYTES and any(uid)
    assert binding.comm_unique_id() != uid
    needed = subprocess.run(["readelf", "-d", str(binding.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libnccl" not in needed


def test_comm_entry_points_reject_bad_arguments():
    """Argument checks of the exchange step that need no GPU."""
    import torch
    from cornelis_b200 import binding
    L = binding.lib()
    assert L.cornelis_cuda_comm_destroy(None) == 0  # like free(NULL)
    assert L.cornelis_cuda_comm_info(None, None, None, None) == binding.ERR_INVALID_ARGUMENT
    assert L.cornelis_cuda_allreduce_framebuffers(None, None, 0) == binding.ERR_INVALID_ARGUMENT
    assert L.cornelis_cuda_reduce_framebuffers(None, 0) == binding.ERR_INVALID_ARGUMENT
    assert L.cornelis_cuda_comm_unique_id(None) == binding.ERR_INVALID_ARGUMENT
    with pytest.raises(ValueError):
        binding.Comm.init_rank(b"short", 0, 1, 0)
    uid = binding.comm_unique_id()
    for rank, n in ((-1, 2), (2, 2), (0, 0)):
        with pytest.raises(binding.CornelisError) as e:
            binding.Comm.init_rank(uid, rank, n, 0)
        assert e.value.code == binding.ERR_INVALID_ARGUMENT
    if not torch.cuda.is_available():  # no device: the communicator of all GPUs has nothing to be made of
        with pytest.raises(binding.CornelisError) as e:
            binding.Comm.init_all([0])
        assert e.value.code == binding.ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """Nothing under cornelis_b200/, include/, the host sources or tools/ may import, load or link oracle/ or read
    /root/reference: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do."""
    offenders = []
    for base in (ROOT / "cornelis_b200", ROOT / "include", ROOT / "tools"):
        for path in base.rglob("*"):
            if path.is_file() and path.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"}:
                text = path.read_text(errors="ignore")
                if re.search(r"\boracle[/.]|from oracle|import oracle|ora_[a-z]+\(|libcornelis_(ref|oracle)|/root/reference", text):
                    offenders.append(str(path.relative_to(ROOT)))
    assert not offenders, offenders


def test_scene_fixtures_are_well_formed():
    from cornelis_b200 import scenes
    c = scenes.cornell_box()
    assert c["spheres"].shape == (4, 4) and c["planes"].shape == (5, 9) and c["materials"].shape == (5, 11)
    m = scenes.microbench_scene(1024)
    assert m["spheres"].shape == (1024, 4) and m["planes"].shape == (6, 9)
    assert (m["sphere_mat"] == np.arange(1024) % 6).all() and (m["plane_mat"] == (1024 + np.arange(6)) % 6).all()
    o, d = scenes.microbench_rays(1000)
    assert np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-6) and np.abs(o).max() <= 1000
