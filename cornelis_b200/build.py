"""Builds the native parts IN-TREE (so the artefacts travel with a gpurun snapshot):

  cornelis_b200/lib/libcornelis_cuda.so   the C-ABI library (include/cornelis_cuda.h): sm_100a kernels + driver
  cornelis_b200/lib/libcorneliscore.so    the C++ host API mirror (include/cornelis/*.hpp) on top of the C-ABI
  cornelis_b200/lib/cornelis              the CLI

nvcc cross-compiles for sm_100a without a GPU.  Flags that matter for parity:
  --fmad=false             the intersection path must reproduce the reference's t bit for bit; the reference is
                           compiled without FMA contraction (hard part 1)
  (default) -prec-div=true -prec-sqrt=true -ftz=false: IEEE division / square root, denormals kept
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "cornelis_b200" / "csrc"
HOST = ROOT / "cornelis_b200" / "host"
LIB = ROOT / "cornelis_b200" / "lib"
OBJ = LIB / "obj"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [*ARCH, "-lineinfo", "-O3", "--fmad=false", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off",
              "-Xptxas", "-v", *os.environ.get("CORNELIS_NVCC_EXTRA", "").split()]


def csrc_hash() -> str:
    """SHA-256 over the device sources (cornelis_b200/csrc, sorted by name) and the nvcc flags: what a committed ncu
    capture or ptxas log has to match to describe the kernels that are running."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(CSRC.iterdir()):
        if path.suffix in {".cu", ".cuh", ".h"}:
            h.update(path.name.encode())
            h.update(path.read_bytes())
    h.update(" ".join(str(f) for f in NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, log: Path | None = None):
    proc = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if log is not None:
        log.write_text(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("build step failed: " + " ".join(str(c) for c in cmd))
    return proc


def build_cuda(force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    headers = [*CSRC.glob("*.cuh"), *CSRC.glob("*.h"), ROOT / "include" / "cornelis_cuda.h", Path(__file__)]
    objs = []
    procs = []
    for src in sorted(CSRC.glob("*.cu")):
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *headers]):
            cmd = [NVCC, *NVCC_FLAGS, "-I", ROOT / "include", "-c", src, "-o", obj]
            procs.append((src, obj, subprocess.Popen([str(c) for c in cmd], stdout=subprocess.PIPE,
                                                    stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        (OBJ / (src.stem + ".ptxas.log")).write_text(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    so = LIB / "libcornelis_cuda.so"
    if force or _stale(so, objs):
        _run([NVCC, *ARCH, "-shared", "-o", so, *objs, "-Xcompiler", "-fPIC", "-ldl"])  # NCCL is dlopen'ed (comm.cu)
    return so


def build_host(force: bool = False):
    """The C++ host mirror of the reference API and the CLI (plain g++, links the C-ABI library)."""
    srcs = sorted(HOST.glob("*.cpp"))
    if not srcs:
        return None
    cxx = os.environ.get("CXX") or "g++"
    inc = ["-I", ROOT / "include"]
    headers = list((ROOT / "include").rglob("*.h*"))
    core_srcs = [s for s in srcs if s.name != "cornelis_cli.cpp"]
    core = LIB / "libcorneliscore.so"
    if force or _stale(core, [*core_srcs, *headers, Path(__file__)]):
        _run([cxx, "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wextra", "-shared", *inc, "-o", core,
              *core_srcs, f"-L{LIB}", "-lcornelis_cuda", "-lz", "-Wl,-rpath,$ORIGIN", "-pthread"])
    cli_src = HOST / "cornelis_cli.cpp"
    cli = LIB / "cornelis"
    if cli_src.exists() and (force or _stale(cli, [cli_src, core, *headers])):
        _run([cxx, "-std=c++17", "-O2", "-Wall", "-Wextra", *inc, "-o", cli, cli_src, f"-L{LIB}", "-lcorneliscore",
              "-lcornelis_cuda", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return core


def build_all(force: bool = False):
    so = build_cuda(force)
    build_host(force)
    return so


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
