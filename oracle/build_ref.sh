#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY: builds oracle/_ref/libusrt_ref.so FROM THE REFERENCE'S OWN TEXT.
#
# The reference's hot path is HLSL compute + a few static C# functions. None of it builds with its own toolchain here
# (no Unity / DXC / .NET), but the kernels are plain C-like code: this recipe feeds the files WHERE THEY LIE under
# $REF (default /root/reference) through a purely syntactic sed pass and compiles them with g++ against two small shim
# headers (oracle/ref_shim/hlsl_shim.hpp, unity_shim.hpp) that supply HLSL / UnityEngine types and intrinsics.
# Outputs (generated sources and the .so) go to oracle/_ref/ only, which is git-ignored: no reference source is
# copied into the repository. The .so travels to the GPU box with the snapshot; it is the pin of oracle/usrt_oracle.cpp
# (tests/test_ref_pin.py) and the generator of tests/golden/ref_*.json (tests/golden/make_ref_golden.py).
#
# What the sed pass changes (syntax only; every arithmetic expression, branch and buffer access is the reference's):
#   HLSL : drop `#pragma ...`, `#include <UnityShaderVariables.cginc>` and `[numthreads(..)]` lines; drop `: SV_*`
#          semantics; `1e-8`, `0.5`, `0.4` -> `1e-8f` ... (an HLSL literal is fp32; a C++ one would be a double);
#          `.xyz` / `.xy` swizzles -> `.xyz()` / `.xy()`; one hook line `USRT_REF_HIT(id, result);` ahead of
#          Raytracing.compute:178 so the hit record can be read out (the reference never stores it).
#          Sorting/*.compute additionally run under a lock-step wave emulator (oracle/ref_shim/wave_emulator.hpp): every
#          thread of a 1024-thread group is a coroutine, group barriers and WavePrefixCountBits / WavePrefixSum block
#          until their participants have arrived.
#   C#   : `private static`/`public` dropped, `out T x` -> `T& x`, `new T(` -> `T(`, `Math.` -> `Math::`,
#          `new AABB { min = min, max = max }` -> `AABB { .min = min, .max = max }`.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
SHADERS="$REF/Assets/_Shaders"
[ -f "$SHADERS/BVH/BVH.compute" ] || { echo "build_ref.sh: no reference checkout at $REF" >&2; exit 3; }
mkdir -p "$OUT"

hlsl() {  # common HLSL -> C++ spelling
    sed -E -e '/^#pragma/d' -e '/#include <UnityShaderVariables.cginc>/d' -e '/^\[numthreads\(/d' \
           -e 's/ : SV_[A-Za-z]+//g' \
           -e 's/\b([0-9]+\.[0-9]+([eE][-+]?[0-9]+)?|[0-9]+[eE][-+]?[0-9]+)\b/\1f/g' \
           -e 's/\.xyz\b/.xyz()/g' -e 's/\.xy\b/.xy()/g' "$1"
}
hlsl "$SHADERS/BVH/BVH.compute" > "$OUT/gen_BVH.inc"
hlsl "$SHADERS/Raytracing/Raytracing.compute" \
    | sed -E 's/^(\s*)const Triangle t = triangleData\[result\.triangleIndex\];/\1USRT_REF_HIT(id, result);\n&/' \
    > "$OUT/gen_Raytracing.inc"
grep -q 'USRT_REF_HIT' "$OUT/gen_Raytracing.inc" || { echo "build_ref.sh: hook line not placed" >&2; exit 4; }

# the five sort kernels (run under oracle/ref_shim/wave_emulator.hpp)
for f in LocalRadixSort Scan GlobalRadixSort; do
    hlsl "$SHADERS/Sorting/$f.compute" > "$OUT/gen_$f.inc"
done

# MeshBufferContainer.cs: ExpandBits .. NormalizeCentroid (:32-83) and DistributeKeys (:154-169)
CS="$REF/Assets/_Scripts/MeshBufferContainer.cs"
sed -n '32,83p;154,169p' "$CS" \
    | sed -E -e 's/private static /static /' -e 's/public void /void /' -e 's/out (Vector3|AABB) /\1\& /g' \
             -e 's/new (Vector3|AABB)\b/\1/g' -e 's/Math\./Math::/g' \
             -e 's/^(\s+)min = min,/\1.min = min,/' -e 's/^(\s+)max = max\s*$/\1.max = max/' \
    > "$OUT/gen_MeshBufferContainer.inc"
grep -q 'static uint ExpandBits' "$OUT/gen_MeshBufferContainer.inc" && grep -q 'void DistributeKeys' "$OUT/gen_MeshBufferContainer.inc" \
    || { echo "build_ref.sh: MeshBufferContainer.cs line ranges moved" >&2; exit 4; }

CXX="${CXX:-g++}"
# same arithmetic flags as the oracle: IEEE fp32 per operation, no contraction, no fast-math, no -march
FLAGS="-O2 -std=c++20 -ffp-contract=off -fno-fast-math -fPIC -pthread -Wno-narrowing -Wno-unknown-pragmas -w -I$REF -I$HERE/ref_shim"
for tu in ref_mesh ref_bvh ref_raytracing ref_sort; do
    $CXX $FLAGS -c "$HERE/ref_shim/$tu.cpp" -o "$OUT/$tu.o"
done
$CXX -shared -pthread -o "$OUT/libusrt_ref.so" "$OUT/ref_mesh.o" "$OUT/ref_bvh.o" "$OUT/ref_raytracing.o" "$OUT/ref_sort.o"
echo "$OUT/libusrt_ref.so"
