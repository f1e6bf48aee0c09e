#!/bin/bash
# A/B of the REF_AABB streaming kernel's shape on the GPU box: rebuilds pp_collide.o with the given macros.
# usage: profiles/sweep_aabb.sh "WARPS CTAS STAGES" ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  set -- $cfg
  rm -f autonomouscar_b200/csrc/build/pp_collide.o
  make -s -C autonomouscar_b200/csrc EXTRA="-DPP_AABB_WARPS=$1 -DPP_AABB_CTAS=$2 -DPP_AABB_STAGES=$3" > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  echo "warps=$1 ctas=$2 stages=$3: $(python profiles/prof_collide.py 24 1 1 20 2>&1 | tail -1)"
done
rm -f autonomouscar_b200/csrc/build/pp_collide.o
make -s -C autonomouscar_b200/csrc > /dev/null 2>&1
