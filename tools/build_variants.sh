#!/bin/bash
# Builds the experimental variants of libptb200.so that profiles/r02_experiments.md compares (into _variants/, git-ignored; they travel to the GPU box).
#   tools/build_variants.sh            then on the box:  tools/sweep_variants_r02.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p _variants
b() { name=$1; shift; python -m pathtracerwithcuda_b200.build --out=$PWD/_variants/$name.so "$@" > /dev/null && echo "built $name: $*"; }
b r01_base    -DPTB_SMEM_STACK=0 -DPTB_SMEM_STACK8=0 -DPTB_LEAF_SINGLE=0 &
b smem16      -DPTB_LEAF_SINGLE=0 &
b leaf1       -DPTB_SMEM_STACK=0 -DPTB_SMEM_STACK8=0 &
b smem24      -DPTB_SMEM_STACK=24 -DPTB_SMEM_STACK8=12 &
wait
b smem32      -DPTB_SMEM_STACK=32 -DPTB_SMEM_STACK8=16 &
b smem12      -DPTB_SMEM_STACK=12 -DPTB_SMEM_STACK8=6 &
b blocks9     -DPTB_PERSISTENT_MIN_BLOCKS=9 &
wait
