"""autonomouscar_b200 -- B200-native hot path of the WPI AutonomousCar local planner.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/prius_planner.h, prius_costmap.h), the ctypes bindings (`planner.Planner`,
`costmap.LocalCostmap`, `costmap.GlobalRoadMap`), the C++ host mirrors of the reference nodes
(`host/`), closed-loop roll-outs (`rollout`) and synthetic-workload helpers for tests and
bench (`worlds`).
"""
from .capi import *  # noqa: F401,F403  (enums, structs, default_params, make_query)
from .results import PlanBatch, PlanRecords, PlanResult  # noqa: F401


def __getattr__(name):
    # the binding needs torch + the built library; import it lazily so that pure-host
    # helpers (capi, worlds) stay importable everywhere
    if name in ("Planner", "PlannerError", "PlannerGroup"):
        from . import planner
        return getattr(planner, name)
    if name in ("LocalCostmap", "GlobalRoadMap"):
        from . import costmap
        return getattr(costmap, name)
    raise AttributeError(name)
