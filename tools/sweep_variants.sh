#!/bin/bash
# usage: tools/sweep_variants.sh "<profile_step args>" variant1.so variant2.so ...
args="$1"; shift
echo "default:"; python tools/profile_step.py $args | tail -1
for v in "$@"; do echo "$v:"; PTB200_LIB=$v python tools/profile_step.py $args | tail -1; done
