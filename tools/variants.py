"""A/B builds of the C-ABI library with different -D switches, for timing kernel variants in one GPU session.

    python tools/variants.py build name1="-DX=1 -DY=0" name2="..."     (here: cross-compiles into cornelis_b200/lib/variants/)
    python tools/variants.py run [bench args...]                       (on the GPU box: bench.py once per variant)
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
VAR = ROOT / "cornelis_b200" / "lib" / "variants"
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17", "-Xcompiler",
         "-fPIC,-ffp-contract=off", "-shared", "-I", str(ROOT / "include")]


def build(specs):
    VAR.mkdir(parents=True, exist_ok=True)
    for old in VAR.glob("*.so"):
        old.unlink()
    procs = []
    for spec in specs:
        name, _, defs = spec.partition("=")
        out = VAR / f"{name}.so"
        srcs = sorted((ROOT / "cornelis_b200" / "csrc").glob("*.cu"))
        procs.append((name, subprocess.Popen([NVCC, *FLAGS, *defs.split(), "-o", str(out), *map(str, srcs)],
                                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out)
            raise SystemExit(f"variant {name} failed to build")
        print("built", name)


def run(args):
    for so in sorted(VAR.glob("*.so")):
        env = dict(os.environ, CORNELIS_CUDA_LIB=str(so))
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--no-cpu-baseline", *args], env=env,
                           capture_output=True, text=True)
        try:
            line = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{so.stem:28s} {line['value']:9.1f} {line['unit']}  e2e {line['e2e']['value']:9.1f}  "
                  f"{line.get('pipeline')}", flush=True)
        except Exception:
            print(so.stem, "FAILED", r.stdout[-300:], r.stderr[-300:], flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
