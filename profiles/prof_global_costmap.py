"""Global road map (row N3) timing: python profiles/prof_global_costmap.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from autonomouscar_b200 import GlobalRoadMap

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
z = np.load(os.path.join(ROOT, "tests", "golden", "road_edges.npz"))
ms = []
for k in range(reps):
    gm = GlobalRoadMap(z["xy"], z["offsets"], device=0, want_host=False)
    ms.append(gm.build_ms)
    gm.close()
print("pg_build device ms:", [round(m, 3) for m in ms], "launches", gm.launches)
