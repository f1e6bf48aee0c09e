/*
 * prius_planner.h -- C ABI of the B200-native Prius local-planner hot path.
 *
 * This is the drop-in boundary for ONE path of the reference repository
 * (WPI-Motion-Planning-Group-1-Fall-2018/AutonomousCar): the inner loop of
 * catkin_ws/src/local_planner/src/local_planner.cpp ("LP" below), i.e.
 * Prius::LocalPlanner::planPath (LP:25-82) and everything it calls, plus the three
 * data-parallel primitives BASELINE.json:north_star names (Philox state sampler,
 * nearest-neighbour over the tree, footprint-vs-occupancy-grid collision).
 *
 * The reference has no FFI of its own: the path sits behind three ROS callbacks of a
 * class whose methods are all private (local_planner.hpp:28-84).  Each entry point below
 * therefore names the reference member function(s) it replaces; INTEGRATION.md shows the
 * ROS-side shim that binds them.
 *
 * Conventions
 *   - every function returns a pp_status (0 = PP_OK); nothing throws across the boundary;
 *   - plain pointers and sizes only; "host" pointers are ordinary CPU memory, "dev"
 *     pointers are CUDA device pointers on the handle's device;
 *   - a handle is single-caller (like the single-threaded ros::spin() of
 *     local_planner_node.cpp:14-17); distinct handles are independent;
 *   - there is NO CPU fallback: if no CUDA device is usable pp_create fails with
 *     PP_ERR_CUDA.
 */
#ifndef PRIUS_PLANNER_H_
#define PRIUS_PLANNER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 1

/* ------------------------------------------------------------------------------------ */
/* status codes                                                                         */
/* ------------------------------------------------------------------------------------ */
typedef enum pp_status {
  PP_OK = 0,
  PP_ERR_INVALID = 1,      /* bad argument (null pointer, negative size, bad enum)        */
  PP_ERR_CUDA = 2,         /* CUDA runtime error; pp_last_error() has the text            */
  PP_ERR_NO_GRID = 3,      /* planning / collision call before pp_set_grid                */
  PP_ERR_CAPACITY = 4,     /* a caller-provided output buffer is too small                */
  PP_ERR_UNSUPPORTED = 5,  /* parameter combination outside what the kernels implement    */
  PP_ERR_NCCL = 6          /* NCCL missing or failed (pp_group_* only)                    */
} pp_status;

/* outcome of one plan query (pp_result.status) */
typedef enum pp_plan_status {
  PP_PLAN_FOUND = 0,       /* LP:43-51: goal test hit, path + profile produced            */
  PP_PLAN_CAP = 1,         /* expansion / node cap reached.  Replaces the 1 s ros::Time
                              guard of LP:33-39; like the reference, nothing is produced  */
  PP_PLAN_EXHAUSTED = 2,   /* frontier ran empty (m_frontier.top() on an empty queue is
                              UB in the reference, LP:40 / LP:106)                        */
  PP_PLAN_BROKEN = 3       /* parent walk of LP:509-527 made no progress in a full pass
                              (the reference would spin forever)                          */
} pp_plan_status;

/* collision semantics */
typedef enum pp_collision_mode {
  PP_COLLISION_AS_SHIPPED = 0, /* LP:445 `return false;` -- what the reference binary does */
  PP_COLLISION_REF_AABB = 1,   /* the dead code LP:446-457 + LP:467-496 enabled verbatim:
                                  axis-aligned cell box at the edge end point, yaw ignored  */
  PP_COLLISION_ORIENTED = 2    /* extension: yaw-rotated footprint sample lattice           */
} pp_collision_mode;

/* how pp_set_grid interprets `data` */
typedef enum pp_grid_layout {
  PP_GRID_OCCUPANCY_MSG = 0,   /* nav_msgs/OccupancyGrid.data as received on /local_costmap;
                                  converted with LP:425-441: cspace[i][j] = data[H*W - i*H - j - 1] */
  PP_GRID_CSPACE = 1           /* already in the planner's m_c_space order: data[i*H + j]   */
} pp_grid_layout;

/* ------------------------------------------------------------------------------------ */
/* parameters: the 11 ROS parameters of LP:390-423 + footprint + caps                    */
/* ------------------------------------------------------------------------------------ */
typedef struct pp_params {
  /* /local_costmap_node/local_costmap_{res,width,height}; LP:394-398 derives the cell
     counts as int(metres / res).  Here the CELL counts are given directly.               */
  double costmap_res;               /* m per cell (launch: 0.25)                          */
  int32_t costmap_width;            /* cells, first index of m_c_space  (launch: 300)     */
  int32_t costmap_height;           /* cells, second index of m_c_space (launch: 300)     */
  /* ~private params, LP:415-422 */
  double collision_buffer_distance; /* read by the reference but never applied (A11)      */
  double time_step_ms;              /* 150 */
  double velocity_res;              /* 0.1 */
  double heading_res;               /* 0.005 */
  double heading_diff;              /* 0.065 */
  double goal_pos_tolerance;        /* 0.1 */
  double goal_speed_tolerance;      /* 0.5 */
  double search_res;                /* 0.25 */
  double max_velocity;              /* m_max_velocity, local_planner.hpp:141 (15)          */
  /* footprint: what setupCollision (LP:328-360) derives from the URDF through tf          */
  double car_length;                /* 4.7   */
  double car_width;                 /* 1.586 */
  int32_t collision_mode;           /* pp_collision_mode                                   */
  /* deterministic replacements of the wall-clock guard LP:33-39 */
  int32_t max_expansions;           /* iterations of the LP:31 loop                        */
  int32_t max_nodes;                /* nodes ever opened (open + closed)                   */
  /* extension: sampling planner (rows S1-S3) */
  int32_t rrt_samples_per_round;    /* K: samples drawn and resolved per synchronous round */
  int32_t rrt_max_samples;          /* total sample budget                                 */
  double rrt_goal_bias;             /* probability that a sample is the goal itself        */
  double rrt_goal_tolerance;        /* |dx|,|dy| box around the goal that ends the search  */
  /* A18: LP:276-283 calls an UNQUALIFIED abs().  With the toolchain the reference mandates (ROS Kinetic =
     Ubuntu 16.04 / GCC 5, whose <cmath> predates the <stdlib.h> overload wrapper) that may resolve to
     int abs(int): the differences are truncated toward zero before the comparison, i.e. the tolerances
     silently become "< 1.0".  0 (default) = the double overload, what g++ 13 and oracle/_ref resolve to;
     1 = the int reading.                                                                              */
  int32_t goal_test_int_abs;
  int32_t reserved[7];
} pp_params;

/* one plan request == one prius_msgs/LocalNav message (already in the base_link frame,
   i.e. after convertGoalToLocalFrame LP:670-691) + the current speed that
   initializePlanner reads from /gazebo/model_states (LP:92). */
typedef struct pp_query {
  double goal_x, goal_y;   /* LocalNav.x / .y                                              */
  double goal_speed;       /* LocalNav.speed                                               */
  double max_speed;        /* LocalNav.max_speed                                           */
  double max_accel;        /* LocalNav.max_accel                                           */
  double start_speed;      /* |prius_velocity.linear|, LP:92                               */
  /* extension fields (sampling planner only; ignored by pp_plan) */
  double start_x, start_y, start_heading;
  uint64_t seed;
  uint32_t query_id;       /* Philox stream id                                             */
  uint32_t reserved;       /* pp_plan_batch_banked: index of this query's grid in the bank; else ignored */
} pp_query;

/* Output of one plan.  The caller sets the cap_* fields and the array pointers (any array
   pointer may be NULL = not wanted); the library fills the n_* fields and the arrays.
   All arrays are HOST memory. */
typedef struct pp_result {
  /* in: capacities, in elements (nodes / path poses / expansions) */
  int32_t cap_nodes;
  int32_t cap_path;
  int32_t cap_expansions;
  /* out */
  int32_t status;          /* pp_plan_status                                               */
  int32_t n_expansions;    /* completed iterations of the LP:31 loop                       */
  int32_t n_nodes;         /* nodes ever opened; node id = order of openNode() calls       */
  int32_t n_open;          /* m_tree_open.size() at return                                 */
  int32_t n_closed;        /* m_tree_closed.size() at return                               */
  int32_t n_path;          /* nav_msgs/Path poses (LP:530-538)                             */
  int32_t n_profile;       /* entries of the MotionPlanning message (LP:548-581)           */
  int32_t n_candidates;    /* children generated by getNeighbors over the whole plan       */
  int32_t n_collisions;    /* of those, rejected by checkCollision (LP:159)                */
  uint64_t tree_digest;    /* sum over slots of a splitmix64 chain seeded with the slot id
                              and fed the bit patterns of the node's 6 key doubles          */
  /* per node, indexed by node id:
       node_state[8*id + 0..7] = child.x child.y parent.x parent.y heading velocity g cost
       node_parent[id]  = first node in scan order (open list, then closed list) whose
                          child_point == this node's parent_point  (LP:511-526)
       node_closed_seq[id] = position in m_tree_closed, or -1 if still open               */
  double* node_state;
  int32_t* node_parent;
  int32_t* node_closed_seq;
  int32_t* expansion_ids;  /* node id expanded at iteration k (LP:40)                      */
  double* path_xy;         /* 2*n_path, start -> goal (LP:530-538)                         */
  int32_t* path_node_ids;  /* node id whose child_point is path pose k                     */
  double* yaw_rates;       /* MotionPlanning.yaw_rates  (LP:576-580)                       */
  double* durations;       /* MotionPlanning.durations, ms                                 */
  double* speeds;          /* MotionPlanning.speeds                                        */
} pp_result;

typedef struct pp_handle pp_handle;   /* one planner instance on one GPU */
typedef struct pp_group pp_group;     /* one handle per device + an NCCL communicator */

/* ------------------------------------------------------------------------------------ */
/* lifetime                                                                             */
/* ------------------------------------------------------------------------------------ */

/* launch-file defaults (full_sim.launch:77-79,90-97; local_planner.hpp:141) */
int pp_default_params(pp_params* out);
int pp_abi_version(void);
/* text of the last error on this thread ("" if none); never NULL */
const char* pp_last_error(void);

/* Replaces LocalPlanner::LocalPlanner (LP:5-18): setupCostmap, setupCSpace,
   getTreeParams, setupCollision. */
int pp_create(const pp_params* params, int device, pp_handle** out);
int pp_destroy(pp_handle* h);
int pp_get_params(const pp_handle* h, pp_params* out);
/* change collision semantics without re-uploading the grid */
int pp_set_collision_mode(pp_handle* h, int mode);

/* ------------------------------------------------------------------------------------ */
/* grid ingest                                                                          */
/* ------------------------------------------------------------------------------------ */

/* Replaces costmapCallback -> mapCostmapToCSpace -> calcGridLocation (LP:658-661,
   LP:425-441).  The caller keeps ownership of `data`; the library copies it to the device,
   applies the layout flip there and builds its own device-side representations.
   width/height must equal the handle's costmap_width/height. */
int pp_set_grid(pp_handle* h, const int8_t* data_host, int32_t width, int32_t height,
                double res, int layout);
/* same, `data_dev` already on the handle's device (no H2D copy) */
int pp_set_grid_dev(pp_handle* h, const int8_t* data_dev, int32_t width, int32_t height,
                    double res, int layout);
/* Scenario sweeps with one costmap per simulated car: a BANK of n grids of the handle's shape, already on
   the device, width*height bytes apart (e.g. what pc_batch_grids_dev returns), built in one set of launches.
   Independent of pp_set_grid: the handle's own grid is untouched. */
int pp_set_grid_bank_dev(pp_handle* h, int32_t n, const int8_t* grids_dev, int32_t width, int32_t height,
                         double res, int layout);
/* read back the planner-order grid (m_c_space, 1 byte per cell) for inspection */
int pp_get_cspace(pp_handle* h, int8_t* out_host, size_t capacity);

/* ------------------------------------------------------------------------------------ */
/* collision: checkCollision + calcCollisionMatrix (LP:443-496), batched                */
/* ------------------------------------------------------------------------------------ */

/* poses = n x (x, y, yaw) float32, array-of-structs, vehicle (base_link) frame.
   out_flags[k] = 1 if pose k collides under the handle's collision mode, else 0.
   Host pointers; H2D / D2H copies happen inside the call. */
int pp_collide_batch(pp_handle* h, int64_t n, const float* poses_host, uint8_t* out_flags_host);
/* Compact host format for PCIe-bound callers (6 instead of 12 bytes per pose): pose k = (qx, qy, qyaw) int16,
   x = x0 + float(qx) * dx,  y = y0 + float(qy) * dy,  yaw = float(qyaw) * dyaw   (float32, each operation rounded
   separately, evaluated on the device exactly like that).  The flags are those of pp_collide_batch on the
   dequantised float32 poses, bit for bit.  Host pointers; H2D / D2H copies happen inside the call. */
typedef struct pp_pose_quant {
  float x0, dx, y0, dy, dyaw;
} pp_pose_quant;
int pp_collide_batch_q16(pp_handle* h, int64_t n, const int16_t* poses_q16_host, const pp_pose_quant* quant,
                         uint8_t* out_flags_host);
/* device pointers; asynchronous on the handle's stream; pp_sync() to wait */
int pp_collide_batch_dev(pp_handle* h, int64_t n, const float* poses_dev, uint8_t* out_flags_dev);
/* pp_collide_batch_dev with ONE BIT per pose as the result: bit b of out_bits_dev[k] = pose 32 k + b, exactly what
   pp_pack_flags_dev makes of pp_collide_batch_dev's flags; ceil(n / 32) words, the unused bits of the last word
   are 0.  In ORIENTED mode the batch kernel writes the bits itself (no byte array, no second pass): the form in
   which results leave the device for a gather.  Asynchronous on the handle's stream. */
int pp_collide_batch_bits_dev(pp_handle* h, int64_t n, const float* poses_dev, uint32_t* out_bits_dev);

/* Result flags as one bit per pose (word k bit b = flag 32k + b; (n + 31) / 32 words, the unused bits of the
   last word are 0): what the multi-GPU result gather moves (NCCL only to gather
   results -- 2 MiB instead of 16 MiB per 2^24 poses).  Asynchronous on the handle's stream; flags_dev must be
   16-byte aligned. */
int pp_pack_flags_dev(pp_handle* h, int64_t n, const uint8_t* flags_dev, uint32_t* out_bits_dev);
int pp_unpack_flags_dev(pp_handle* h, int64_t n, const uint32_t* bits_dev, uint8_t* out_flags_dev);
/* along-edge variant (row S3): edge k runs from (x0,y0) to (x1,y1) with constant yaw; the
   footprint is tested every search_res/2 (the spacing rule of LP:308).
   edges = n x (x0, y0, x1, y1, yaw) float32 */
int pp_collide_edges_dev(pp_handle* h, int64_t n, const float* edges_dev, uint8_t* out_flags_dev);
/* debug / parity: the transformed footprint sample lattice of ONE pose, world metres.
   out_xy has room for 2*capacity doubles; *n_points receives the lattice size. */
int pp_footprint_points(pp_handle* h, double x, double y, double yaw, double* out_xy_host,
                        int32_t capacity, int32_t* n_points);

/* ------------------------------------------------------------------------------------ */
/* sampler (row S1) and nearest neighbour (rows A6 / S2)                                */
/* ------------------------------------------------------------------------------------ */

/* Philox4x32-10 state sampler: pose k = f(seed, stream, first_index + k), uniform over
   x in [x_lo,x_hi), y in [y_lo,y_hi), yaw in [-pi,pi).  Writes n x 3 float32. */
int pp_sample_poses_dev(pp_handle* h, uint64_t seed, uint32_t stream, uint64_t first_index,
                        int64_t n, float x_lo, float x_hi, float y_lo, float y_hi,
                        float* out_poses_dev);
/* raw generator, for known-answer tests: out[4] = philox4x32_10(counter[4], key[2]) */
int pp_philox4x32_10(pp_handle* h, const uint32_t counter[4], const uint32_t key[2],
                     uint32_t out[4]);

/* brute-force nearest neighbour: for each of n_query points the index of the nearest of
   n_nodes points (squared Euclidean distance in float32, ties -> lowest index).
   node_xy / query_xy are float2 arrays (device). */
int pp_nearest_dev(pp_handle* h, int64_t n_nodes, const float* node_xy_dev, int64_t n_query,
                   const float* query_xy_dev, int32_t* out_index_dev, float* out_dist2_dev);
/* exact-match mode == the std::find + GraphNode::operator== lookups of LP:56-57 /
   graph_node.hpp:24-30: keys are 6 doubles (child.x child.y parent.x parent.y heading
   velocity); out_index = first node equal to the query key, or -1. */
int pp_find_exact_dev(pp_handle* h, int64_t n_nodes, const double* node_keys_dev,
                      int64_t n_query, const double* query_keys_dev, int32_t* out_index_dev);

/* ------------------------------------------------------------------------------------ */
/* planning                                                                             */
/* ------------------------------------------------------------------------------------ */

/* Replaces localNavCallback -> planPath (LP:663-668, LP:25-82) incl. calcPathMsg and
   calcMPMessage (LP:503-583).  Synchronous, like the callback. */
int pp_plan(pp_handle* h, const pp_query* query, pp_result* result);
/* n independent queries against the same grid (BASELINE config 5) */
int pp_plan_batch(pp_handle* h, int32_t n, const pp_query* queries, pp_result* results);
/* n independent queries, query k against grid queries[k].reserved of the bank (pp_set_grid_bank_dev):
   N simulated cars, each planning on its own costmap, in one launch */
int pp_plan_batch_banked(pp_handle* h, int32_t n, const pp_query* queries, pp_result* results);
/* Device-resident results ("fixed-stride result records gathered via NCCL").
   Query k leaves one record of pp_plan_record_stride(cap_path) bytes at out_records_dev + k * stride:

     pp_plan_record_header (64 B) | path_xy[cap_path][2] f64 | yaw_rates[cap_path] f64 | durations[cap_path] f64 |
     speeds[cap_path] f64 | path_node_ids[cap_path] i32        (unused entries: 0 / -1; padded to 16 B)

   i.e. everything planPath publishes (LP:503-583) plus the tree summary.  The call is ASYNCHRONOUS on the
   handle's stream: nothing is copied back, so the records can go straight into an NCCL all-gather
   (autonomouscar_b200/sharding.py, pp_group_plan_batch) and only the rank that wants them copies them to the
   host, once.  `queries_host` is read before the call returns when it is pageable memory; pinned memory must
   stay valid until the stream reaches the kernel.  A path longer than cap_path sets `truncated`. */
typedef struct pp_plan_record_header {
  int32_t status, n_expansions, n_nodes, n_open, n_closed, n_path, n_profile, n_candidates, n_collisions;
  int32_t truncated;       /* 1: n_path > cap_path, the arrays hold the first cap_path poses                    */
  int32_t cap_path;        /* the array length this record was written with                                     */
  int32_t reserved;
  uint64_t tree_digest;
  uint64_t reserved2;
} pp_plan_record_header;
size_t pp_plan_record_stride(int32_t cap_path);
int pp_plan_batch_dev(pp_handle* h, int32_t n, const pp_query* queries_host, int32_t cap_path, void* out_records_dev);
/* host-side helper: scatter n records (HOST memory, written with `cap_path`) into pp_result structs whose
   cap_path / array pointers the caller has set, like pp_plan_batch would have filled them */
int pp_plan_records_unpack(const void* records_host, int32_t n, int32_t cap_path, pp_result* results);
/* Debug bitmap of the most recent single-query pp_plan: out[i*H + j] = 1 for every cell that
   markVisited (LP:318-326) touched -- one mark per child returned by getNeighbors, cleared at
   the start of each plan like clearVisited (LP:251-262).  The reference never reads it on the
   path (checkVisited, LP:239, is unused); unlike the reference the write is bounds-guarded. */
int pp_get_visited(pp_handle* h, uint8_t* out_host, size_t capacity);
/* extension: round-synchronous sampling planner (sample -> nearest -> steer -> edge
   collision -> append).  Uses the query's start_*, seed and query_id fields. */
int pp_rrt_plan(pp_handle* h, const pp_query* query, pp_result* result);
int pp_rrt_plan_batch(pp_handle* h, int32_t n, const pp_query* queries, pp_result* results);

/* ------------------------------------------------------------------------------------ */
/* streams, timing, accounting                                                          */
/* ------------------------------------------------------------------------------------ */
int pp_sync(pp_handle* h);
/* cudaStream_t of the handle, as an opaque pointer (for event timing by the caller) */
int pp_stream(pp_handle* h, void** out_stream);
/* number of kernels this handle has launched since creation */
int pp_launch_count(const pp_handle* h, uint64_t* out);
/* device time (ms, CUDA events on the handle's stream) of the phases of the most recent
   collide / plan call: out_ms[0..n).  Phase ids are PP_PHASE_*. */
enum { PP_PHASE_H2D = 0, PP_PHASE_BROAD = 1, PP_PHASE_NARROW = 2, PP_PHASE_D2H = 3,
       PP_PHASE_PLAN = 4, PP_PHASE_COUNT = 8 };
int pp_last_phase_ms(pp_handle* h, float* out_ms, int32_t n);
/* units (poses / queries) the narrow-phase kernel of the last collide call processed */
int pp_last_narrow_count(pp_handle* h, int64_t* out);
/* Roofline helper: random 32-byte-sector gathers (16 B loaded per sector, 32 B counted) from a
   buffer of `buffer_bytes` (rounded down to a power of two); best of 4 timed launches, GB/s.
   With a buffer that fits L2 this is the L2 gather bandwidth the collision kernels are bounded
   by when they read one sector per footprint cell. */
int pp_bench_l2_gather(pp_handle* h, size_t buffer_bytes, double* out_gbps);
/* tuning knob: 0 = every pose goes through the full footprint gather ("direct"),
   1 = conservative broad phase first (default).  Results are identical. */
int pp_set_broad_phase(pp_handle* h, int enabled);

/* ------------------------------------------------------------------------------------ */
/* multi-GPU, single process (row 8(e)): one handle per device, NCCL only to gather      */
/* ------------------------------------------------------------------------------------ */
int pp_group_create(const pp_params* params, int n_devices, const int* devices,
                    pp_group** out);
int pp_group_destroy(pp_group* g);
int pp_group_size(const pp_group* g, int* out);
int pp_group_set_grid(pp_group* g, const int8_t* data_host, int32_t width, int32_t height,
                      double res, int layout);
/* poses sharded contiguously: device d gets [d*n/G, (d+1)*n/G); flags gathered with
   ncclAllGather so every device (and the host copy) holds all n flags in input order */
int pp_group_collide_batch(pp_group* g, int64_t n, const float* poses_host,
                           uint8_t* out_flags_host);
/* queries sharded contiguously by index; every device leaves fixed-stride records (pp_plan_batch_dev) in its send
   buffer, ONE ncclAllGather moves them over NVLink, the host copy comes from device 0.  (Results that ask for node
   arrays / expansion ids are planned per device through pp_plan_batch instead, without a gather.) */
int pp_group_plan_batch(pp_group* g, int32_t n, const pp_query* queries, pp_result* results);
/* device `device_index`'s copy of the records gathered by the last pp_group_plan_batch, input order */
int pp_group_records_dev(pp_group* g, int device_index, const void** out_records_dev);

#ifdef __cplusplus
}
#endif
#endif /* PRIUS_PLANNER_H_ */
