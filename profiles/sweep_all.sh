#!/bin/bash
# A/B of build parameters that live in shared headers: rebuilds the whole library with the given -D flags.
# usage: profiles/sweep_all.sh "-DPP_NEAR_BINS=16" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  rm -f autonomouscar_b200/csrc/build/*.o
  make -s -j8 -C autonomouscar_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "$flags: $(python profiles/prof_collide.py 24 1 2 10 2>&1 | tail -1)"
done
rm -f autonomouscar_b200/csrc/build/*.o
make -s -j8 -C autonomouscar_b200/csrc > /dev/null 2>&1
