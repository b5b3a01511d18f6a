#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the reference itself as the strong oracle / CPU baseline.
#
# Compiles the UNMODIFIED sources under /root/reference (where they lie; nothing is copied into this repository)
# plus oracle/ref_harness.cpp into oracle/_ref/libcornelis_ref.so.  oracle/_ref/ is git-ignored but NOT
# gpurun-ignored, so the built library travels to the GPU box, where /root/reference does not exist.
#
# The reference's own CMake build cannot configure here (TBB, fmt, xsimd, Catch2 are neither installed nor
# fetchable), so the path's few source files are compiled directly:
#   * -I oracle/shims first: header-only stand-ins for tbb / loguru / fmt (scheduling and logging only);
#   * a generated shadow of include/cornelis/Math.hpp in which `floatN<N>` is spelled `floatN` on the five
#     defaulted special members (lines 190-193, 198) — g++ 13 rejects the injected-class-name-with-arguments
#     form that clang accepts; no arithmetic is touched.  The shadow is produced by sed into oracle/_ref/ at
#     build time and never committed;
#   * -include algorithm (Color.cpp uses std::clamp without the header);
#   * -O2, no -march/-mfma, -ffp-contract=off: the reference's CMake sets no optimisation or arch flags and a
#     fused multiply-add would change the bits of t.
set -euo pipefail
REF="${CORNELIS_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: $REF not present — keeping any prebuilt $OUT/libcornelis_ref.so" >&2
    exit 0
fi
mkdir -p "$OUT/shadow/cornelis" "$OUT/obj"
sed -e '190,193s/floatN<N>/floatN/g' -e '198s/floatN<N>/floatN/g' "$REF/include/cornelis/Math.hpp" \
    > "$OUT/shadow/cornelis/Math.hpp"

CXX="${CXX:-g++}"
FLAGS=(-std=c++20 -O2 -w -fPIC -ffp-contract=off -pthread
       -I"$HERE/shims" -I"$OUT/shadow" -I"$REF/include" -I"$REF/src" -isystem "$REF/external"
       -include algorithm -include cstdlib)
pids=()
for unit in Geometry Materials Camera Scene Color Tiles NanoVDBMath; do
    "$CXX" "${FLAGS[@]}" -c "$REF/src/$unit.cpp" -o "$OUT/obj/$unit.o" &
    pids+=($!)
done
"$CXX" "${FLAGS[@]}" -I"$HERE" -c "$HERE/ref_harness.cpp" -o "$OUT/obj/ref_harness.o" &
pids+=($!)
"$CXX" -O2 -w -fPIC -I"$REF/src/extern" -c "$REF/src/extern/stb_image_write.cpp" -o "$OUT/obj/stb_image_write.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$CXX" -shared -pthread -o "$OUT/libcornelis_ref.so" "$OUT"/obj/*.o
echo "built $OUT/libcornelis_ref.so"
