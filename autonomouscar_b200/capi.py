"""ctypes mirror of include/prius_planner.h (structs, enums, default parameters).

No computation happens here.  The structs are shared by the product binding
(`autonomouscar_b200.planner`) and by the test-only oracle binding (`oracle/binding.py`),
so that parity tests hand the same bytes to both sides.
"""
import ctypes as C

PP_ABI_VERSION = 1

# pp_status
PP_OK, PP_ERR_INVALID, PP_ERR_CUDA, PP_ERR_NO_GRID, PP_ERR_CAPACITY, PP_ERR_UNSUPPORTED, PP_ERR_NCCL = range(7)
# pp_plan_status
PP_PLAN_FOUND, PP_PLAN_CAP, PP_PLAN_EXHAUSTED, PP_PLAN_BROKEN = range(4)
# pp_collision_mode
PP_COLLISION_AS_SHIPPED, PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED = range(3)
# pp_grid_layout
PP_GRID_OCCUPANCY_MSG, PP_GRID_CSPACE = range(2)
# phases of pp_last_phase_ms
PP_PHASE_H2D, PP_PHASE_BROAD, PP_PHASE_NARROW, PP_PHASE_D2H, PP_PHASE_PLAN = range(5)
PP_PHASE_COUNT = 8


class pp_params(C.Structure):
    _fields_ = [
        ("costmap_res", C.c_double),
        ("costmap_width", C.c_int32),
        ("costmap_height", C.c_int32),
        ("collision_buffer_distance", C.c_double),
        ("time_step_ms", C.c_double),
        ("velocity_res", C.c_double),
        ("heading_res", C.c_double),
        ("heading_diff", C.c_double),
        ("goal_pos_tolerance", C.c_double),
        ("goal_speed_tolerance", C.c_double),
        ("search_res", C.c_double),
        ("max_velocity", C.c_double),
        ("car_length", C.c_double),
        ("car_width", C.c_double),
        ("collision_mode", C.c_int32),
        ("max_expansions", C.c_int32),
        ("max_nodes", C.c_int32),
        ("rrt_samples_per_round", C.c_int32),
        ("rrt_max_samples", C.c_int32),
        ("rrt_goal_bias", C.c_double),
        ("rrt_goal_tolerance", C.c_double),
        ("goal_test_int_abs", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class pp_pose_quant(C.Structure):
    """pp_collide_batch_q16: x = x0 + float(qx) * dx, y = y0 + float(qy) * dy, yaw = float(qyaw) * dyaw (float32)."""
    _fields_ = [("x0", C.c_float), ("dx", C.c_float), ("y0", C.c_float), ("dy", C.c_float), ("dyaw", C.c_float)]


def quantise_poses(poses, x_lo, x_hi, y_lo, y_hi):
    """float32 (x, y, yaw) -> (int16 triples, pp_pose_quant, the float32 poses the device will reconstruct)."""
    import numpy as np
    poses = np.asarray(poses, dtype=np.float32).reshape(-1, 3)
    k = pp_pose_quant()
    k.dx, k.dy = np.float32((x_hi - x_lo) / 65535.0), np.float32((y_hi - y_lo) / 65535.0)
    k.x0, k.y0 = np.float32(x_lo + 32768.0 * float(k.dx)), np.float32(y_lo + 32768.0 * float(k.dy))
    k.dyaw = np.float32(np.pi / 32768.0)
    q = np.empty(poses.shape, dtype=np.int16)
    q[:, 0] = np.clip(np.rint((poses[:, 0] - np.float32(k.x0)) / np.float32(k.dx)), -32768, 32767)
    q[:, 1] = np.clip(np.rint((poses[:, 1] - np.float32(k.y0)) / np.float32(k.dy)), -32768, 32767)
    q[:, 2] = np.clip(np.rint(poses[:, 2] / np.float32(k.dyaw)), -32768, 32767)
    back = np.empty(poses.shape, dtype=np.float32)
    back[:, 0] = np.float32(k.x0) + q[:, 0].astype(np.float32) * np.float32(k.dx)
    back[:, 1] = np.float32(k.y0) + q[:, 1].astype(np.float32) * np.float32(k.dy)
    back[:, 2] = q[:, 2].astype(np.float32) * np.float32(k.dyaw)
    return q, k, back


class pp_query(C.Structure):
    _fields_ = [
        ("goal_x", C.c_double),
        ("goal_y", C.c_double),
        ("goal_speed", C.c_double),
        ("max_speed", C.c_double),
        ("max_accel", C.c_double),
        ("start_speed", C.c_double),
        ("start_x", C.c_double),
        ("start_y", C.c_double),
        ("start_heading", C.c_double),
        ("seed", C.c_uint64),
        ("query_id", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class pp_result(C.Structure):
    _fields_ = [
        ("cap_nodes", C.c_int32),
        ("cap_path", C.c_int32),
        ("cap_expansions", C.c_int32),
        ("status", C.c_int32),
        ("n_expansions", C.c_int32),
        ("n_nodes", C.c_int32),
        ("n_open", C.c_int32),
        ("n_closed", C.c_int32),
        ("n_path", C.c_int32),
        ("n_profile", C.c_int32),
        ("n_candidates", C.c_int32),
        ("n_collisions", C.c_int32),
        ("tree_digest", C.c_uint64),
        ("node_state", C.POINTER(C.c_double)),
        ("node_parent", C.POINTER(C.c_int32)),
        ("node_closed_seq", C.POINTER(C.c_int32)),
        ("expansion_ids", C.POINTER(C.c_int32)),
        ("path_xy", C.POINTER(C.c_double)),
        ("path_node_ids", C.POINTER(C.c_int32)),
        ("yaw_rates", C.POINTER(C.c_double)),
        ("durations", C.POINTER(C.c_double)),
        ("speeds", C.POINTER(C.c_double)),
    ]


def default_params(**overrides):
    """Launch-file values: full_sim.launch:77-79,90-97 and local_planner.hpp:141."""
    p = pp_params()
    p.costmap_res = 0.25
    p.costmap_width = 300
    p.costmap_height = 300
    p.collision_buffer_distance = 1.0
    p.time_step_ms = 150.0
    p.velocity_res = 0.1
    p.heading_res = 0.005
    p.heading_diff = 0.065
    p.goal_pos_tolerance = 0.1
    p.goal_speed_tolerance = 0.5
    p.search_res = 0.25
    p.max_velocity = 15.0
    p.car_length = 4.7
    p.car_width = 1.586
    p.collision_mode = PP_COLLISION_AS_SHIPPED
    p.max_expansions = 100000
    p.max_nodes = 60000      # ~ what the reference reaches within its 1 s guard (LP:33-39)
    p.rrt_samples_per_round = 256
    p.rrt_max_samples = 1 << 20
    p.rrt_goal_bias = 0.05
    p.rrt_goal_tolerance = 1.0
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def make_query(goal_x, goal_y, goal_speed=1.5, max_speed=1.5, max_accel=1.5, start_speed=0.0,
               start_x=0.0, start_y=0.0, start_heading=0.0, seed=0, query_id=0, grid_index=0):
    """LocalNav defaults from local_navigator.py:524-527.  grid_index: which grid of the planner's bank this
    query plans on (pp_plan_batch_banked only)."""
    q = pp_query()
    q.reserved = grid_index
    q.goal_x, q.goal_y, q.goal_speed = goal_x, goal_y, goal_speed
    q.max_speed, q.max_accel, q.start_speed = max_speed, max_accel, start_speed
    q.start_x, q.start_y, q.start_heading = start_x, start_y, start_heading
    q.seed, q.query_id = seed, query_id
    return q
