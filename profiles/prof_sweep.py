"""One tick of the batched sweep (256 cars, own costmaps, ORIENTED): python profiles/prof_sweep.py [cars] [ticks]
pc_build_batch -> pp_set_grid_bank_dev -> pp_plan_batch_banked; same workload as bench.py's bench_sweep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cars = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
print(bench.bench_sweep(None, cars=cars, ticks=ticks))
