/*
 * prius_costmap.h -- C ABI of the B200-native local costmap builder.
 *
 * The step directly BEFORE the planner's hot path: the reference node
 *   catkin_ws/src/local_costmap/src/local_costmap.cpp ("LC")
 * turns two planar laser scans, one 16 x 512 block-laser cloud and the global road map into the
 * 300 x 300 int8 occupancy grid the planner consumes (/local_costmap).  Built on the device, the grid
 * never crosses PCIe: pc_grid_dev() hands the device pointer straight to pp_set_grid_dev()
 * (prius_planner.h) with layout PP_GRID_OCCUPANCY_MSG.
 *
 * The reference has no FFI; the boundary is the node contract (LC:5-29, 360-385).  tf look-ups of
 * the reference become plain numbers in pc_scan_tf / pc_inputs (the ROS shim reads them from tf).
 *
 * Conventions as in prius_planner.h: every function returns a pp_status, nothing throws, plain
 * pointers and sizes, a handle is single-caller, no CPU fallback.
 */
#ifndef PRIUS_COSTMAP_H_
#define PRIUS_COSTMAP_H_

#include <stddef.h>
#include <stdint.h>

#include "prius_planner.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ROS parameters of the node (LC:14-26, full_sim.launch:71-72,77-82).  The reference keeps the cell
   counts as DOUBLES (metres / res, LC:20-21,25-26) and so does this struct. */
typedef struct pc_params {
  double res;                   /* local_costmap_res (0.25)                                    */
  double height_cells;          /* local_costmap_height / res (300)                            */
  double width_cells;           /* local_costmap_width  / res (300)                            */
  double z_ground_buffer;       /* 0.075                                                       */
  double obstacle_range;        /* 2                                                           */
  double max_inflation_radius;  /* 5                                                           */
  double global_height_cells;   /* /global_costmap_node/global_costmap_height / res (12000)    */
  double global_width_cells;    /* /global_costmap_node/global_costmap_width  / res (12000)    */
  int32_t reserved[8];
} pc_params;

/* One sensor_msgs/LaserScan (front_right / front_left planar laser, LC:372-380) plus the numbers
   mapPlanarScan reads from tf (LC:84-108, 127, 141-142). */
typedef struct pc_planar_scan {
  const float* ranges;          /* host pointer, n entries (NULL / n = 0: no scan received yet) */
  int32_t n;
  float angle_min, angle_increment, range_max;
  double roll;                  /* getRPY of tf(frame <- center_laser_link), LC:108              */
  double tf_x, tf_y;            /* origin of tf(frame <- center_laser_link), LC:117-124          */
  double z_map;                 /* origin.z of tf(map <- frame), LC:127                          */
  double inv_x, inv_y;          /* origin of the inverse of tf(frame <- center_laser_link), LC:141 */
} pc_planar_scan;

/* Everything one centerLaserCallback -> pubOccGrid -> calcOccGrid pass reads (LC:71-80, 366-370) */
typedef struct pc_inputs {
  pc_planar_scan right, left;   /* LC:75-76 */
  const float* cloud_xyz;       /* sensor_msgs/PointCloud.points, n_cloud x (x, y, z) float32, host */
  int32_t n_cloud;
  double center_z;              /* origin.z of tf(center_laser_link <- map), LC:164, 176          */
  /* tf(map <- base_link) and its inverse as 3 x 4 row-major [R | t] (LC:217, 244, 248); only used
     when a global grid has been set */
  double map_from_base[12];
  double base_from_map[12];
} pc_inputs;

typedef struct pc_handle pc_handle;

/* launch-file defaults */
int pc_default_params(pc_params* out);
/* Replaces LocalCostmap::LocalCostmap + setupCostmap (LC:5-68) */
int pc_create(const pc_params* params, int device, pc_handle** out);
int pc_destroy(pc_handle* h);
/* Replaces globalCostmapCallback (LC:382-385): the library keeps a device copy (144 MB at the
   launch-file size).  len may be 0 (mapRoadVectors then returns at once, LC:210-213). */
int pc_set_global_grid(pc_handle* h, const int8_t* data_host, int64_t len);
/* Replaces centerLaserCallback -> pubOccGrid -> calcOccGrid (LC:71-80, 360-370): clearMap,
   mapPlanarScan x 2, mapCenterScan, inflateGrid, mapRoadVectors.  out_host (width*height int8, the
   nav_msgs/OccupancyGrid.data of /local_costmap) may be NULL when only the device grid is wanted. */
int pc_build(pc_handle* h, const pc_inputs* in, int8_t* out_host);
/* n independent scenes (scenario sweeps: one per simulated car) through the same kernels in ONE set of
   launches -- a scene index rides on blockIdx.y.  Results are identical to n pc_build calls.  out_host
   (n * width*height int8, scene-major) and out_masks (n region masks as pc_front_regions returns them)
   may each be NULL; the grids stay on the device (pc_batch_grids_dev) until the next batch.  The global
   grid, if set, is shared by all scenes.  Does not disturb the pc_build grid. */
int pc_build_batch(pc_handle* h, int32_t n, const pc_inputs* in, int8_t* out_host, uint32_t* out_masks);
int pc_batch_grids_dev(pc_handle* h, const int8_t** out_dev, int32_t* n_scenes);
/* device pointer of the most recent grid (message order), valid until the next pc_build */
int pc_grid_dev(pc_handle* h, const int8_t** out_dev, int32_t* width, int32_t* height);
/* cudaStream_t of the handle; kernels launched since creation; device ms of the last pc_build */
int pc_stream(pc_handle* h, void** out_stream);
int pc_launch_count(const pc_handle* h, uint64_t* out);
int pc_last_build_ms(pc_handle* h, float* out_ms);

/* Row N4: the region scan of the query generator, local_navigator.py:81-257 (callbackOccGrid), on the
   most recent device grid -- the costmap does not cross PCIe for it either.  The navigator reads
   board[x][y] = data[x * info.width + y] and looks at x 120..179, y 180..299 only:
     F_k (x 140..159), k = 1..4 for y 180.., 210.., 240.., 270..: ANY cell == 100
     R_k (x 120..139), L_k (x 160..179): the verdict of the LAST cell scanned (their else-branches)
     obstacle: self.obstacle once the scan is over = (any F_k) and no L cell == 100 (the L regions are
               scanned last and their hits write False, :214-257) */
#define PC_REGION_F(k) (1u << ((k) - 1))          /* k = 1..4 */
#define PC_REGION_R(k) (1u << (4 + (k) - 1))
#define PC_REGION_L(k) (1u << (8 + (k) - 1))
#define PC_REGION_OBSTACLE (1u << 12)
#define PC_REGION_ANY_L_HIT (1u << 13)
#define PC_REGION_R_SCANNED(k) (1u << (16 + (k) - 1)) /* region non-empty at this grid size: flag was rewritten */
#define PC_REGION_L_SCANNED(k) (1u << (20 + (k) - 1))
int pc_front_regions(pc_handle* h, uint32_t* out_mask);
/* the same scan over any device-resident OccupancyGrid.data (len cells, info.width = info_width), on the
   handle's stream */
int pc_front_regions_dev(pc_handle* h, const int8_t* grid_dev, int32_t info_width, int64_t len, uint32_t* out_mask);

/* ------------------------------------------------------------------------------------ */
/* row N3: the static global road map (catkin_ws/src/global_costmap/src/global_costmap.cpp, */
/* "GC"): road polylines -> the 12000 x 12000 lane-cost grid published once on            */
/* /global_costmap and consumed by mapRoadVectors (LC:208-257).                           */
/* ------------------------------------------------------------------------------------ */
typedef struct pg_params {
  double res;            /* /local_costmap_node/local_costmap_res (GC:8), 0.25            */
  double road_width;     /* ~road_width, 11.1 (full_sim.launch:69)                        */
  int32_t num_lanes;     /* ~num_lanes, 2 (:70)                                           */
  int32_t reserved0;
  double height_cells;   /* global_costmap_height / res (GC:14), 12000                    */
  double width_cells;    /* global_costmap_width  / res (GC:15), 12000                    */
  int32_t reserved[8];
} pg_params;

int pg_default_params(pg_params* out);
/* Replaces GlobalCostmap::GlobalCostmap -> setupRoadMap (GC:5-18, 52-61) minus the file parsing
   (GC:25-50 is the caller's; the +5 / -5 shift of GC:42-43 IS applied here): polyline k =
   points [offsets[k], offsets[k+1]) of xy_host (float32 x, y pairs as read from edges.csv).
   The grid stays on `device`; out_host (height*width int8, message order) may be NULL. */
int pg_build(const pg_params* params, int device, const float* xy_host, const int32_t* offsets_host,
             int32_t n_sections, int8_t* out_host, int8_t** out_dev);
/* device time (ms) of the most recent pg_build on this thread, and its kernel launch count */
int pg_last_build(float* out_ms, int32_t* out_launches);
/* hand a pg_build device grid to a costmap builder without copying (it takes ownership) */
int pc_adopt_global_grid(pc_handle* h, int8_t* grid_dev, int64_t len);
int pg_free(int device, int8_t* grid_dev);

#ifdef __cplusplus
}
#endif
#endif /* PRIUS_COSTMAP_H_ */
