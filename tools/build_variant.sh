#!/bin/bash
# Build a compile-time variant of libcq.so next to the default one (no GPU needed):
#   tools/build_variant.sh <tag> "<extra nvcc flags>"      ->  swift-game-engine_b200/csrc/libcq_<tag>.so
# The variant travels to the GPU box with gpurun (CQ_LIB=.../libcq_<tag>.so selects it; tools/ab_variants.sh).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TAG=$1; FLAGS=$2
B=/tmp/cq_variant_$TAG
rm -rf $B; mkdir -p $B
cp $ROOT/swift-game-engine_b200/csrc/{Makefile,*.cu,*.cuh,*.h,*.cpp} $B/
sed -i "s#\.\./\.\./include#$ROOT/include#g" $B/Makefile $B/*.cu $B/*.cuh $B/*.h $B/*.cpp
make -C $B -j8 -s EXTRA="$FLAGS"
cp $B/libcq.so $ROOT/swift-game-engine_b200/csrc/libcq_$TAG.so
echo "built libcq_$TAG.so ($FLAGS)"
