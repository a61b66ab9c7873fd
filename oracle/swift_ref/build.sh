#!/bin/bash
# build.sh — golden-vector harness: compile the reference's own Swift sources, UNMODIFIED, against the simd shim.
#
#   oracle/swift_ref/build.sh <path to the reference checkout>       ->  oracle/_ref/cq_swift_ref
#
# Needs swiftc (5.9+; Linux or macOS).  There is none in this repository's build image, so this script has never run
# here: it is the one-command route to pin the C++ oracle (oracle/cq_oracle.cpp) and the CUDA path to the reference's own
# code wherever a Swift toolchain exists:
#
#   python oracle/swift_ref/export_inputs.py /tmp/cq_inputs.bin          # scenes + seeded queries of C1 / C2 / C3-sample
#   SWIFT_DETERMINISTIC_HASHING=1 oracle/_ref/cq_swift_ref /tmp/cq_inputs.bin /tmp/cq_outputs.bin
#   python oracle/swift_ref/import_goldens.py /tmp/cq_inputs.bin /tmp/cq_outputs.bin   # -> tests/golden/swift_*.npz
#   python -m pytest tests/test_swift_goldens.py                         # oracle (CPU) and library (GPU) against them
#
# Nothing of the reference is copied into this repository: the sources are read where they lie.  Game/Systems.swift is
# compiled as THREE unmodified line ranges (the file also holds animation / render-extract systems that need Metal-side
# types): :8-22 (protocols, isActive), :205-250 (PhysicsIntentSystem) + :410-435 (approachVec, d3, f3),
# :596-2210 (GravitySystem ... AgentSeparationSystem — everything SURVEY.md §8a cites).  If the upstream file has moved,
# adjust RANGES; the ranges must start and end on top-level declaration boundaries.
set -euo pipefail
REF=${1:?usage: build.sh <reference checkout>}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
B="$OUT/swift_build"
RANGES='8,22p;205,250p;410,435p;596,2210p'
mkdir -p "$B"
# 1. the shim as a module named `simd`
swiftc -O -parse-as-library -emit-module -emit-library -module-name simd -o "$B/libsimd.so" \
       -emit-module-path "$B/simd.swiftmodule" "$HERE/simd_shim.swift"
# 2. the physics part of Systems.swift (unmodified lines, concatenated)
sed -n "$RANGES" "$REF/Game/Systems.swift" > "$B/Systems_physics.swift"
# 3. everything together; -enable-testing is not needed: the harness only uses public API
swiftc -O -I "$B" -L "$B" -lsimd -Xlinker -rpath -Xlinker "$B" -o "$OUT/cq_swift_ref" \
       "$REF/Game/Entity.swift" "$REF/Game/World.swift" "$REF/Game/Components.swift" "$REF/Game/ProceduralMeshAPI.swift" \
       "$REF/Game/CollisionQuery.swift" "$B/Systems_physics.swift" "$HERE/Stubs.swift" "$HERE/main.swift"
echo "built $OUT/cq_swift_ref"
