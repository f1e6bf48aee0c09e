#!/bin/bash
# A/B of the persistent sampling-planner kernel on the GPU box: rebuilds pp_rrt.o with the given -D flags.
# usage: profiles/sweep_rrt.sh "-DPP_RRT_GRID=64" "-DPP_RRT_GRID=64 -DPP_RRT_WARP_CELLS=0" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  rm -f autonomouscar_b200/csrc/build/pp_rrt.o
  make -s -C autonomouscar_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "$flags: $(python profiles/prof_rrt_config4.py 256 2>&1 | grep 'rep 1' | sed 's/.*wall/wall/') | full budget: $(python profiles/prof_rrt_config4.py 256 1048576 1000 0 2>&1 | grep 'rep 1' | sed 's/.*wall/wall/')"
done
rm -f autonomouscar_b200/csrc/build/pp_rrt.o
make -s -C autonomouscar_b200/csrc > /dev/null 2>&1
