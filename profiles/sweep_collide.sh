#!/bin/bash
# A/B of the ORIENTED batch kernel's build parameters on the GPU box: rebuilds pp_collide.o with the given -D flags.
# usage: profiles/sweep_collide.sh "-DPP_COLLIDE_PPT=4 -DPP_COLLIDE_MIN_BLOCKS=4" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  rm -f autonomouscar_b200/csrc/build/pp_collide.o
  make -s -C autonomouscar_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "$flags: $(python profiles/prof_collide.py 24 1 2 10 2>&1 | tail -1)"
done
rm -f autonomouscar_b200/csrc/build/pp_collide.o
make -s -C autonomouscar_b200/csrc > /dev/null 2>&1
