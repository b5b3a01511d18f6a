"""Opcode histogram of named kernels in the built library, from `cuobjdump -sass` (static SASS, no GPU needed):
what the committed evidence for FFMA2 / FMNMX3 / REDG.E.ADD.F32x4 and for the absence of tensor-core and TMA opcodes
is made of.

    python tools/sass_histogram.py [kernel-substring ...] > profiles/r2_sass/opcodes.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cornelis_b200 import build  # noqa: E402

LIB = ROOT / "cornelis_b200" / "lib" / "libcornelis_cuda.so"
wanted = sys.argv[1:] or ["k_persistent_queuedILb0", "k_intersect_batchILb0", "k_walk", "k_shadeILb0", "k_intersectILb0",
                          "k_resolve_srgb8", "k_accumulate"]
text = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
head = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}   git {head}   csrc sha256 {build.csrc_hash()}")
arch = re.search(r"arch = (sm_\w+)", text)
print(f"# {arch.group(0) if arch else ''}")
functions = re.split(r"\n\s*Function : ", text)[1:]
INTEREST = ("FFMA2", "FMNMX3", "RED", "ATOM", "HMMA", "IMMA", "UTMA", "UTCMMA", "TCGEN", "LDGSTS", "UBLKCP", "MUFU", "DFMA",
            "VOTE", "LDS", "STS", "LDL", "STL")
for fn in functions:
    name = fn.split("\n", 1)[0].strip()
    if not any(w in name for w in wanted):
        continue
    ops = collections.Counter()
    full = collections.Counter()
    for line in fn.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m:
            full[m.group(1)] += 1
            ops[m.group(1).split(".")[0]] += 1
    total = sum(ops.values())
    print(f"\n== {name}\n   {total} instructions ({total * 16 / 1024:.1f} KB), {len(ops)} distinct opcodes")
    print("   " + "  ".join(f"{op} {n}" for op, n in ops.most_common(24)))
    notable = {k: v for k, v in full.items() if any(k.startswith(p) for p in INTEREST)}
    print("   notable: " + "  ".join(f"{op} {n}" for op, n in sorted(notable.items(), key=lambda kv: -kv[1])[:30]))
    tensor = sum(v for k, v in full.items() if re.match(r"(HMMA|IMMA|DMMA|QMMA|UTCMMA|UTMA|TCGEN|UBLKCP)", k))
    print(f"   tensor-core / TMA opcodes: {tensor}")
