#!/bin/bash
# A/B of the ORIENTED batch kernel's build parameters on the GPU box: rebuilds the objects that hold the collision
# code with the given -D flags.
# usage: profiles/sweep_collide.sh "-DPP_COLLIDE_MIN_BLOCKS=3 -DPP_COARSE4_UNROLL=4" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  rm -f autonomouscar_b200/csrc/build/pp_collide.o
  make -s -C autonomouscar_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "$flags: $(python profiles/prof_collide.py 24 1 2 10 2>&1 | tail -1)"
done
rm -f autonomouscar_b200/csrc/build/pp_collide.o
make -s -C autonomouscar_b200/csrc > /dev/null 2>&1
