#!/bin/bash
# On the GPU box: every variant of tools/build_variants.sh through tools/sweep_tree.py on one workload (bench schedule: 16 passes x 4 streams).
#   tools/sweep_variants_r02.sh c2 [variant ...]
wl=${1:-c2}; shift
variants=${@:-"r01_base smem16 leaf1 DEFAULT smem12 smem24 smem32 blocks9"}
export PIF=${PIF:-16} STREAMS=${STREAMS:-4}
for v in $variants; do
  if [ "$v" = DEFAULT ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/_variants/$v.so; fi
  printf "%-10s " $v; python tools/sweep_tree.py $wl "" 2>&1 | tail -1
done
