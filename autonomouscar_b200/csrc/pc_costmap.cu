// pc_costmap.cu -- the local costmap builder on the device (include/prius_costmap.h, row N2).
//
// One calcOccGrid pass of catkin_ws/src/local_costmap/src/local_costmap.cpp ("LC"):
//   clearMap LC:349-358 . mapPlanarScan x2 LC:82-156 . mapCenterScan LC:158-206 .
//   inflateGrid LC:288-347 . mapRoadVectors LC:208-257 . calcGridLocation LC:259-265 .
//   calcGlobalGridLocation LC:267-275 . calcCartesianCoords LC:277-286
//
// Why this parallelises exactly.  The reference writes its grid sequentially, but every write is
// one of three monotone rules: an obstacle mark (100) always sticks, a free mark (0) never
// overwrites 100 (LC:136, 149, 183, 199), unknown is -1; inflateGrid applies "reached by an
// occupied cell" (100) after "reached by a free cell" (0) (LC:312-330); mapRoadVectors keeps the
// larger of the cell and the global road cost (LC:250-254).  All of them are max-reductions over
// {-1 < 0 < 1..99 < 100}, so the pass is five order-independent atomicMax sweeps over an int32
// working grid -- bit-identical to the sequential result whatever the thread schedule.
//
// Arithmetic: the reference's doubles, operation for operation, through the individually rounded
// helpers of pp_math.cuh (the library is built with -fmad=false); sin/cos/atan2 are the
// deterministic ones shared with the oracle.
//
// Launch structure: inputs are packed into one pinned block (header | right ranges | left ranges |
// cloud) and copied with ONE cudaMemcpyAsync; the memset and five kernels have fixed arguments and fixed grids
// (grid-stride loops over counts read from the device header), so the whole pass is one CUDA graph
// launch.  Work is tiny (9.5 k rays, 90 k cells, 360 k road samples): the bound is launch latency
// and FP64 issue, not memory.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/prius_costmap.h"
#include "pp_common.cuh"
#include "pp_math.cuh"

namespace pp {

struct PcDevHeader {
  pc_params P;
  // scans: 0 = right, 1 = left
  int n_ray[2];
  int ray_off[2];            // float offsets into the payload
  float angle_min[2], angle_increment[2], range_max[2];
  double roll[2], tf_x[2], tf_y[2], z_map[2], inv_x[2], inv_y[2];
  int n_cloud;
  int cloud_off;
  double center_z;
  double map_from_base[12];
  double base_from_map[12];
  long long global_len;      // 0 = no global grid (LC:210-213)
};

struct PcGeom {              // derived once per handle from pc_params, host side, plain double ops
  int len;                   // int(H * W), LC:63
  int h_int;                 // int(H), the divisor of LC:279
  int nx, ny;                // int(H * 2), int(W * 2), LC:235-236
};

// Every kernel takes a scene index from blockIdx.y (pc_build: one scene; pc_build_batch: n of them in
// the same launches).  Scene s reads header s, shares the payload block, and owns work / out slice s.
struct PcScenes {
  const unsigned char* hdrs;   // n headers, kHdrBytes apart
  const float* payload;        // all scenes' floats; the headers hold offsets into it
  int* work;                   // n x (occ | infl), 2 * len int32 each
  int8_t* out;                 // n x len int8 (message order)
  size_t hdr_stride;           // bytes
};
__device__ __forceinline__ const PcDevHeader& pc_hdr(const PcScenes& S) {
  return *reinterpret_cast<const PcDevHeader*>(S.hdrs + blockIdx.y * S.hdr_stride);
}
__device__ __forceinline__ int* pc_occ(const PcScenes& S, int len) { return S.work + static_cast<size_t>(blockIdx.y) * 2 * len; }

__device__ __forceinline__ bool pc_grid_location(const pc_params& P, double x, double y, int* loc) {
  int wp, hp;
  if (!dm::trunc_to_int(dm::add(dm::div(x, P.res), dm::div(P.width_cells, 2.0)), &wp)) return false;   // LC:261
  if (!dm::trunc_to_int(dm::add(dm::div(y, P.res), dm::div(P.height_cells, 2.0)), &hp)) return false;  // LC:262
  return dm::trunc_to_int(dm::sub(dm::add(static_cast<double>(wp), dm::mul(static_cast<double>(hp), P.width_cells)), 1.0),
                          loc);                                                                         // LC:263
}

// ---- mapPlanarScan + mapCenterScan: one warp per ray / point --------------------------------
__global__ void __launch_bounds__(256)
pc_mark_kernel(PcScenes S, int len) {
  const PcDevHeader& H = pc_hdr(S);
  const pc_params& P = H.P;
  const float* __restrict__ payload = S.payload;
  int* __restrict__ occ = pc_occ(S, len);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int total = H.n_ray[0] + H.n_ray[1] + H.n_cloud;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += warps) {
    double x, y, ca, sa;
    int n = 0;
    bool obstacle, clear_ray = true;
    if (item < H.n_ray[0] + H.n_ray[1]) {
      const int s = item < H.n_ray[0] ? 0 : 1;
      const int i = item - (s ? H.n_ray[0] : 0);
      const float r = payload[H.ray_off[s] + i];
      if (r > H.range_max[s]) continue;                                                   // LC:111
      const float anglef = __fadd_rn(H.angle_min[s], __fmul_rn(static_cast<float>(i), H.angle_increment[s]));  // LC:113
      double sn, cs, sr, cr;
      dm::sincos(static_cast<double>(anglef), &sn, &cs);
      dm::sincos(H.roll[s], &sr, &cr);
      const double rd = static_cast<double>(r);
      const double A = dm::mul(dm::mul(rd, cs), cr), B = dm::mul(dm::mul(rd, sn), cr);
      if (s == 0) { y = -dm::sub(A, H.tf_x[0]); x = dm::sub(B, H.tf_y[0]); }              // LC:117-118
      else { y = dm::sub(A, H.tf_x[1]); x = -dm::sub(B, H.tf_y[1]); }                     // LC:123-124
      const double max_range = fabs(dm::div(H.z_map[s], dm::div(sr, cr)));                // LC:127
      obstacle = fabs(y) < max_range && rd > P.obstacle_range;                            // LC:129
      const double dx = dm::sub(x, H.inv_x[s]), dy = dm::sub(y, H.inv_y[s]);
      const double atb = dm::atan2_(dy, dx);                                              // LC:141
      const double dist = sqrt(dm::add(dm::mul(dx, dx), dm::mul(dy, dy)));                // LC:142
      if (!dm::trunc_to_int(dm::div(dist, P.res), &n)) n = 0;                             // LC:143
      dm::sincos(atb, &sa, &ca);
    } else {
      const float* pt = payload + H.cloud_off + 3 * (item - H.n_ray[0] - H.n_ray[1]);
      const float pxf = pt[0], pyf = pt[1], pzf = pt[2];
      x = static_cast<double>(pxf); y = static_cast<double>(pyf);
      const double dist = sqrt(dm::add(dm::mul(x, x), dm::mul(y, y)));                    // LC:175
      const double thresh = dm::add(H.center_z, P.z_ground_buffer);
      obstacle = static_cast<double>(pzf) > thresh && dist > P.obstacle_range;            // LC:176
      const double atb = dm::atan2_(y, x);                                                // LC:188
      if (!dm::trunc_to_int(dm::div(dist, P.res), &n)) n = 0;                             // LC:190
      clear_ray = (pzf == 0.0f ? 1.0 : 0.0) > thresh;                                     // LC:194: `!pt.z > ...`
      dm::sincos(atb, &sa, &ca);
    }
    if (lane == 0) {
      int loc;
      if (pc_grid_location(P, x, y, &loc) && loc >= 0 && loc < len) atomicMax(&occ[loc], obstacle ? 100 : 0);
    }
    if (!clear_ray) continue;
    for (int k = 1 + lane; k < n; k += 32) {                                              // LC:144-153, 192-203
      const double step = dm::mul(static_cast<double>(k), P.res);
      const double x_ = dm::sub(x, dm::mul(step, ca)), y_ = dm::sub(y, dm::mul(step, sa));
      int loc;
      if (pc_grid_location(P, x_, y_, &loc) && loc >= 0 && loc < len) atomicMax(&occ[loc], 0);
    }
  }
}

// ---- inflateGrid (LC:288-325): one thread per source cell ------------------------------------
__global__ void __launch_bounds__(256)
pc_inflate_kernel(PcScenes S, PcGeom G) {
  const pc_params& P = pc_hdr(S).P;
  const int* __restrict__ occ = pc_occ(S, G.len);
  int* __restrict__ infl = pc_occ(S, G.len) + G.len;
  const double max_dist = sqrt(dm::add(dm::mul(dm::mul(P.height_cells, P.res), dm::mul(P.height_cells, P.res)),
                                       dm::mul(dm::mul(P.width_cells, P.res), dm::mul(P.width_cells, P.res))));  // LC:298
  for (int count = blockIdx.x * blockDim.x + threadIdx.x; count < G.len; count += gridDim.x * blockDim.x) {
    const int pt = occ[count];
    if (pt != 0 && pt != 100) continue;        // only free / occupied cells write anything (LC:312-320)
    const int hp = count / G.h_int, wp = count % G.h_int;                                  // LC:279-281
    const double x = dm::mul(dm::sub(static_cast<double>(wp), dm::div(P.width_cells, 2.0)), P.res);
    const double y = dm::mul(dm::sub(static_cast<double>(hp), dm::div(P.height_cells, 2.0)), P.res);
    const double d = sqrt(dm::add(dm::mul(x, x), dm::mul(y, y)));                          // LC:297
    const double q = dm::div(d, max_dist);
    const double rad = dm::mul(dm::mul(q, q), P.max_inflation_radius);                     // LC:299
    const double num = dm::div(rad, P.res);                                                // LC:300
    const double angle_inc = dm::div(6.283185307179586, num);                              // 2 * M_PI / num
    for (int i = 0; static_cast<double>(i) < num; ++i) {
      double s, c;
      dm::sincos(dm::mul(static_cast<double>(i), angle_inc), &s, &c);
      for (int j = 0; static_cast<double>(j) < num; ++j) {
        const double t = dm::mul(dm::div(static_cast<double>(j), num), rad);
        const double x_ = dm::add(x, dm::mul(t, c)), y_ = dm::add(y, dm::mul(t, s));      // LC:307-308
        int loc;
        if (!pc_grid_location(P, x_, y_, &loc)) continue;
        if (loc >= 0 && loc < G.len) atomicMax(&infl[loc], pt);                            // LC:310-320, 327-330
      }
    }
  }
}

// ---- LC:332-346: inflated 0 / 100 replace the cell --------------------------------------------
__global__ void pc_merge_kernel(PcScenes S, int len) {
  int* __restrict__ occ = pc_occ(S, len);
  const int* __restrict__ infl = occ + len;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += gridDim.x * blockDim.x) {
    const int v = infl[k];
    if (v == 100 || v == 0) occ[k] = v;
  }
}

__device__ __forceinline__ void pc_apply(const double* m, double x, double y, double* ox, double* oy) {
  // tf::Transform * Vector3(x, y, 0): row . v + origin, left to right
  *ox = dm::add(dm::add(dm::add(dm::mul(m[0], x), dm::mul(m[1], y)), dm::mul(m[2], 0.0)), m[3]);
  *oy = dm::add(dm::add(dm::add(dm::mul(m[4], x), dm::mul(m[5], y)), dm::mul(m[6], 0.0)), m[7]);
}

// ---- mapRoadVectors (LC:208-257): one thread per sample of the 2H x 2W lattice -----------------
__global__ void __launch_bounds__(256)
pc_roads_kernel(PcScenes S, const int8_t* __restrict__ global_grid, PcGeom G) {
  const PcDevHeader& H = pc_hdr(S);
  int* __restrict__ occ = pc_occ(S, G.len);
  const pc_params& P = H.P;
  if (H.global_len < 1) return;                                                            // LC:210-213
  const double total_x = dm::mul(P.height_cells, P.res), total_y = dm::mul(P.width_cells, P.res);
  const double min_x = dm::div(-total_x, 2.0), min_y = dm::div(-total_y, 2.0);
  const long long total = static_cast<long long>(G.nx) * G.ny;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t / G.ny), j = static_cast<int>(t % G.ny);
    const double ux = dm::add(min_x, dm::mul(dm::div(static_cast<double>(i), static_cast<double>(G.nx)), total_x));  // LC:242
    const double uy = dm::add(min_y, dm::mul(dm::div(static_cast<double>(j), static_cast<double>(G.ny)), total_y));  // LC:243
    double gx, gy, lx, ly;
    pc_apply(H.map_from_base, ux, uy, &gx, &gy);                                           // LC:244
    int cx, cy, wpos, hpos, gloc;                                                          // LC:269-273
    if (!dm::trunc_to_int(dm::sub(dm::div(P.global_height_cells, 2.0), dm::div(gx, P.res)), &cx)) continue;
    if (!dm::trunc_to_int(dm::sub(dm::div(P.global_width_cells, 2.0), dm::div(gy, P.res)), &cy)) continue;
    if (!dm::trunc_to_int(dm::sub(P.global_width_cells, static_cast<double>(cy)), &wpos)) continue;
    if (!dm::trunc_to_int(dm::sub(P.global_height_cells, static_cast<double>(cx)), &hpos)) continue;
    if (!dm::trunc_to_int(dm::add(static_cast<double>(hpos), dm::mul(static_cast<double>(wpos), P.global_height_cells)), &gloc)) continue;
    if (gloc < 0 || gloc >= H.global_len) continue;
    pc_apply(H.base_from_map, gx, gy, &lx, &ly);                                           // LC:248
    int lloc;
    if (!pc_grid_location(P, lx, ly, &lloc) || lloc < 0 || lloc >= G.len) continue;
    atomicMax(&occ[lloc], static_cast<int>(__ldg(&global_grid[gloc])));                    // LC:250-254
  }
}

__global__ void pc_pack_kernel(PcScenes S, int len) {
  const int* __restrict__ occ = pc_occ(S, len);
  int8_t* __restrict__ out = S.out + static_cast<size_t>(blockIdx.y) * len;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += gridDim.x * blockDim.x)
    out[k] = static_cast<int8_t>(occ[k]);
}

// Row N4 (query generator): the region scan of local_navigator.py:112-257 on the device grid.  The
// navigator unpacks the message as board[x][y] = data[x * info.width + y] (:101-110) and walks
// x ascending, y from 100 up; only x in 120..179, y in 180..299 is ever looked at.  Per 30-cell band
// k (F1..F4 at y 180.., 210.., 240.., 270..): F_k = ANY cell == 100 in x 140..159; R_k / L_k keep the
// verdict of the LAST cell scanned in x 120..139 / 160..179 (their else-branches reset the flag).
// Bit layout of the mask: PC_REGION_* in prius_costmap.h.
// One CTA per grid (grids `stride` cells apart).
__global__ void pc_front_regions_kernel(const int8_t* __restrict__ grids, size_t stride, int W, unsigned* __restrict__ out) {
  const int8_t* __restrict__ grid = grids + blockIdx.x * stride;
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0;
  __syncthreads();
  unsigned m = 0;
  for (int t = threadIdx.x; t < 60 * 120; t += blockDim.x) {
    const int x = 120 + t / 120, y = 180 + t % 120;
    if (x >= W || y >= W) continue;
    const bool hit = grid[static_cast<long long>(x) * W + y] == 100;
    const int band = (y - 180) / 30, col = (x - 120) / 20;
    if (col == 1) {
      if (hit) m |= 1u << band;
    } else {
      const int x_last = min(col == 0 ? 139 : 179, W - 1), y_last = min(209 + 30 * band, W - 1);
      if (col == 2 && hit) m |= PC_REGION_ANY_L_HIT;
      if (x == x_last && y == y_last) m |= (hit ? 1u : 0u) << ((col == 0 ? 4 : 8) + band) | 1u << ((col == 0 ? 16 : 20) + band);
    }
  }
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_mask, m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned r = s_mask;
    if ((r & 0xFu) && !(r & PC_REGION_ANY_L_HIT)) r |= PC_REGION_OBSTACLE;   // self.obstacle once the scan is over
    out[blockIdx.x] = r;
  }
}

}  // namespace pp

using namespace pp;

struct pc_handle {
  pc_params P;
  PcGeom G;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int* d_work = nullptr;            // occ | infl, 2 * len int32
  int8_t* d_out = nullptr;          // len int8 (message order)
  int8_t* h_out = nullptr;          // pinned mirror
  int8_t* d_global = nullptr;
  long long global_len = 0;
  unsigned* d_regions = nullptr;    // pc_front_regions result
  unsigned char* h_in = nullptr;    // pinned: PcDevHeader | payload floats
  unsigned char* d_in = nullptr;
  size_t in_cap = 0;                // bytes
  cudaGraphExec_t graph = nullptr;
  uint64_t launches = 0;
  float last_ms = 0;
  // pc_build_batch: n scenes at once (buffers grow on demand)
  int batch_cap = 0, batch_n = 0;
  int* d_work_b = nullptr;          // batch_cap x 2 * len int32
  int8_t* d_out_b = nullptr;        // batch_cap x len int8
  unsigned* d_regions_b = nullptr;  // batch_cap masks
  unsigned char* h_in_b = nullptr;  // pinned: n headers | payloads
  unsigned char* d_in_b = nullptr;
  size_t in_cap_b = 0;
};

namespace {

constexpr size_t kHdrBytes = (sizeof(PcDevHeader) + 255) & ~size_t(255);

int ensure_input(pc_handle* h, size_t bytes) {
  if (bytes <= h->in_cap) return PP_OK;
  size_t cap = 1 << 20;
  while (cap < bytes) cap <<= 1;
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }   // d_in moves: re-capture
  cudaFreeHost(h->h_in); cudaFree(h->d_in);
  h->h_in = nullptr; h->d_in = nullptr; h->in_cap = 0;
  PP_CUDA(cudaMallocHost(&h->h_in, cap));
  PP_CUDA(cudaMalloc(&h->d_in, cap));
  h->in_cap = cap;
  return PP_OK;
}

// the six launches of one pass over n scenes, on `s` (the single-scene pass is captured into a graph)
void launch_pass(pc_handle* h, cudaStream_t s, const PcScenes& S, int n) {
  const int len = h->G.len;
  const int T = 256;
  // a lone scene spreads over the whole GPU; a batch gives every scene a slice and lets the scenes fill it
  const int per_scene = n == 1 ? h->sm_count * 8 : std::max(4, h->sm_count * 16 / n);
  const unsigned cell_blocks = static_cast<unsigned>(std::min((len + T - 1) / T, per_scene));
  const unsigned ny = static_cast<unsigned>(n);
  cudaMemsetAsync(S.work, 0xFF, sizeof(int) * 2 * static_cast<size_t>(len) * n, s);              // clearMap: -1
  pc_mark_kernel<<<dim3(static_cast<unsigned>(per_scene), ny), T, 0, s>>>(S, len);
  pc_inflate_kernel<<<dim3(cell_blocks, ny), T, 0, s>>>(S, h->G);
  pc_merge_kernel<<<dim3(cell_blocks, ny), T, 0, s>>>(S, len);
  if (h->d_global) {
    const long long samples = static_cast<long long>(h->G.nx) * h->G.ny;
    const unsigned road_blocks = static_cast<unsigned>(std::min<long long>((samples + T - 1) / T, 2 * per_scene));
    pc_roads_kernel<<<dim3(road_blocks, ny), T, 0, s>>>(S, h->d_global, h->G);
  }
  pc_pack_kernel<<<dim3(cell_blocks, ny), T, 0, s>>>(S, len);
}

PcScenes single_scene(const pc_handle* h) {
  return PcScenes{h->d_in, reinterpret_cast<const float*>(h->d_in + kHdrBytes), h->d_work, h->d_out, kHdrBytes};
}

// number of payload floats of one scene (+ argument checks shared by pc_build / pc_build_batch)
int scene_floats(const pc_inputs* in, const char* who, size_t* out) {
  const pc_planar_scan* scans[2] = {&in->right, &in->left};
  size_t n = 0;
  for (int s = 0; s < 2; ++s) {
    const int nr = scans[s]->ranges ? scans[s]->n : 0;
    if (nr < 0 || nr > (1 << 22)) { set_last_error(std::string(who) + ": bad scan size"); return PP_ERR_INVALID; }
    n += static_cast<size_t>(nr);
  }
  const int nc = in->cloud_xyz ? in->n_cloud : 0;
  if (nc < 0 || nc > (1 << 24)) { set_last_error(std::string(who) + ": bad cloud size"); return PP_ERR_INVALID; }
  *out = n + 3 * static_cast<size_t>(nc);
  return PP_OK;
}

// header + floats of one scene into the pinned block; `off` = this scene's first float in `payload`
void pack_scene(const pc_handle* h, const pc_inputs* in, PcDevHeader* hd, float* payload, int off) {
  const pc_planar_scan* scans[2] = {&in->right, &in->left};
  hd->P = h->P;
  for (int s = 0; s < 2; ++s) {
    const int nr = scans[s]->ranges ? scans[s]->n : 0;
    hd->n_ray[s] = nr;
    hd->ray_off[s] = off;
    if (nr) std::memcpy(payload + off, scans[s]->ranges, sizeof(float) * nr);
    off += nr;
    hd->angle_min[s] = scans[s]->angle_min; hd->angle_increment[s] = scans[s]->angle_increment;
    hd->range_max[s] = scans[s]->range_max;
    hd->roll[s] = scans[s]->roll; hd->tf_x[s] = scans[s]->tf_x; hd->tf_y[s] = scans[s]->tf_y;
    hd->z_map[s] = scans[s]->z_map; hd->inv_x[s] = scans[s]->inv_x; hd->inv_y[s] = scans[s]->inv_y;
  }
  const int nc = in->cloud_xyz ? in->n_cloud : 0;
  hd->n_cloud = nc;
  hd->cloud_off = off;
  if (nc) std::memcpy(payload + off, in->cloud_xyz, sizeof(float) * 3 * static_cast<size_t>(nc));
  hd->center_z = in->center_z;
  std::memcpy(hd->map_from_base, in->map_from_base, sizeof(hd->map_from_base));
  std::memcpy(hd->base_from_map, in->base_from_map, sizeof(hd->base_from_map));
  hd->global_len = h->d_global ? h->global_len : 0;
}

int capture(pc_handle* h) {
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
  cudaGraph_t g = nullptr;
  PP_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  launch_pass(h, h->stream, single_scene(h), 1);
  PP_CUDA(cudaStreamEndCapture(h->stream, &g));
  const cudaError_t e = cudaGraphInstantiate(&h->graph, g, 0);
  cudaGraphDestroy(g);
  PP_CUDA(e);
  return PP_OK;
}

}  // namespace

extern "C" {

int pc_default_params(pc_params* p) {
  if (!p) return PP_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->res = 0.25;                       // full_sim.launch:77
  p->height_cells = 75 / 0.25;         // :78, LC:20
  p->width_cells = 75 / 0.25;          // :79, LC:21
  p->z_ground_buffer = 0.075;          // :80
  p->obstacle_range = 2;               // :81
  p->max_inflation_radius = 5;         // :82
  p->global_height_cells = 3000 / 0.25;  // :71, LC:25
  p->global_width_cells = 3000 / 0.25;   // :72, LC:26
  return PP_OK;
}

int pc_create(const pc_params* params, int device, pc_handle** out) {
  if (!params || !out) { set_last_error("pc_create: null argument"); return PP_ERR_INVALID; }
  *out = nullptr;
  PP_REQUIRE(params->res > 0, "pc_create: res must be > 0");
  PP_REQUIRE(params->height_cells >= 1 && params->width_cells >= 1, "pc_create: empty grid");
  PP_REQUIRE(params->height_cells <= 16384 && params->width_cells <= 16384, "pc_create: grid too large");
  PP_REQUIRE(params->global_height_cells >= 0 && params->global_width_cells >= 0 &&
             params->global_height_cells * params->global_width_cells < 1073741824.0, "pc_create: global grid too large");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    set_last_error(std::string("pc_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    return PP_ERR_CUDA;
  }
  PP_REQUIRE(device >= 0 && device < n_dev, "pc_create: bad device index");
  PP_CUDA(cudaSetDevice(device));
  pc_handle* h = new pc_handle();
  h->P = *params;
  h->device = device;
  h->G.len = static_cast<int>(params->height_cells * params->width_cells);   // LC:63
  h->G.h_int = static_cast<int>(params->height_cells);                        // LC:279
  h->G.nx = static_cast<int>(params->height_cells * 2);                       // LC:235
  h->G.ny = static_cast<int>(params->width_cells * 2);                        // LC:236
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  auto fail = [&](int rc) { pc_destroy(h); return rc; };
#define PC_TRY(expr) do { if ((expr) != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(cudaGetLastError())); return fail(PP_ERR_CUDA); } } while (0)
  PC_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  PC_TRY(cudaEventCreate(&h->ev0));
  PC_TRY(cudaEventCreate(&h->ev1));
  PC_TRY(cudaMalloc(&h->d_work, sizeof(int) * 2 * static_cast<size_t>(h->G.len)));
  PC_TRY(cudaMalloc(&h->d_out, static_cast<size_t>(h->G.len)));
  PC_TRY(cudaMallocHost(&h->h_out, static_cast<size_t>(h->G.len)));
  PC_TRY(cudaMalloc(&h->d_regions, sizeof(unsigned)));
#undef PC_TRY
  *out = h;
  return PP_OK;
}

int pc_destroy(pc_handle* h) {
  if (!h) return PP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->graph) cudaGraphExecDestroy(h->graph);
  cudaFree(h->d_work); cudaFree(h->d_out); cudaFree(h->d_global); cudaFree(h->d_in); cudaFree(h->d_regions);
  cudaFree(h->d_work_b); cudaFree(h->d_out_b); cudaFree(h->d_regions_b); cudaFree(h->d_in_b);
  cudaFreeHost(h->h_out); cudaFreeHost(h->h_in); cudaFreeHost(h->h_in_b);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return PP_OK;
}

int pc_set_global_grid(pc_handle* h, const int8_t* data_host, int64_t len) {
  if (!h) { set_last_error("pc_set_global_grid: null handle"); return PP_ERR_INVALID; }
  PP_REQUIRE(len >= 0 && (len == 0 || data_host), "pc_set_global_grid: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
  cudaFree(h->d_global);
  h->d_global = nullptr;
  h->global_len = 0;
  if (len > 0) {
    PP_CUDA(cudaMalloc(&h->d_global, static_cast<size_t>(len)));
    PP_CUDA(cudaMemcpyAsync(h->d_global, data_host, static_cast<size_t>(len), cudaMemcpyHostToDevice, h->stream));
    PP_CUDA(cudaStreamSynchronize(h->stream));
    h->global_len = len;
  }
  return PP_OK;
}

int pc_adopt_global_grid(pc_handle* h, int8_t* grid_dev, int64_t len) {
  if (!h) { set_last_error("pc_adopt_global_grid: null handle"); return PP_ERR_INVALID; }
  PP_REQUIRE(len >= 0 && (len == 0 || grid_dev), "pc_adopt_global_grid: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
  cudaFree(h->d_global);
  h->d_global = len > 0 ? grid_dev : nullptr;
  h->global_len = len;
  return PP_OK;
}

int pc_build(pc_handle* h, const pc_inputs* in, int8_t* out_host) {
  if (!h || !in) { set_last_error("pc_build: null argument"); return PP_ERR_INVALID; }
  size_t n_float = 0;
  int rc = scene_floats(in, "pc_build", &n_float);
  if (rc) return rc;
  PP_CUDA(cudaSetDevice(h->device));
  const size_t bytes = kHdrBytes + sizeof(float) * n_float;
  rc = ensure_input(h, bytes);
  if (rc) return rc;
  pack_scene(h, in, reinterpret_cast<PcDevHeader*>(h->h_in), reinterpret_cast<float*>(h->h_in + kHdrBytes), 0);

  if (!h->graph) { rc = capture(h); if (rc) return rc; }
  PP_CUDA(cudaEventRecord(h->ev0, h->stream));
  PP_CUDA(cudaMemcpyAsync(h->d_in, h->h_in, bytes, cudaMemcpyHostToDevice, h->stream));
  PP_CUDA(cudaGraphLaunch(h->graph, h->stream));
  h->launches += h->d_global ? 5 : 4;
  PP_CUDA(cudaEventRecord(h->ev1, h->stream));
  if (out_host) PP_CUDA(cudaMemcpyAsync(h->h_out, h->d_out, static_cast<size_t>(h->G.len), cudaMemcpyDeviceToHost, h->stream));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  if (out_host) std::memcpy(out_host, h->h_out, static_cast<size_t>(h->G.len));
  cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  return PP_OK;
}

int pc_build_batch(pc_handle* h, int32_t n, const pc_inputs* in, int8_t* out_host, uint32_t* out_masks) {
  if (!h || (n > 0 && !in)) { set_last_error("pc_build_batch: null argument"); return PP_ERR_INVALID; }
  PP_REQUIRE(n >= 0 && n <= 65535, "pc_build_batch: 0 <= n <= 65535 scenes");
  // (the region scan reads info.width^2 cells per grid; a grid with height < width has no defined answer)
  PP_REQUIRE(!out_masks || h->P.height_cells >= h->P.width_cells, "pc_build_batch: masks need height >= width (see pc_front_regions)");
  h->batch_n = 0;
  if (n == 0) return PP_OK;
  const size_t len = static_cast<size_t>(h->G.len);
  std::vector<size_t> off(static_cast<size_t>(n) + 1, 0);
  for (int k = 0; k < n; ++k) {
    size_t nf = 0;
    const int rc = scene_floats(&in[k], "pc_build_batch", &nf);
    if (rc) return rc;
    off[k + 1] = off[k] + nf;
  }
  PP_REQUIRE(off[n] < (size_t(1) << 31), "pc_build_batch: more than 2^31 input floats");
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  if (n > h->batch_cap) {
    cudaFree(h->d_work_b); cudaFree(h->d_out_b); cudaFree(h->d_regions_b);
    h->d_work_b = nullptr; h->d_out_b = nullptr; h->d_regions_b = nullptr; h->batch_cap = 0;
    PP_CUDA(cudaMalloc(&h->d_work_b, sizeof(int) * 2 * len * n));
    PP_CUDA(cudaMalloc(&h->d_out_b, len * n));
    PP_CUDA(cudaMalloc(&h->d_regions_b, sizeof(unsigned) * n));
    h->batch_cap = n;
  }
  const size_t hdr_bytes = kHdrBytes * static_cast<size_t>(n);
  const size_t bytes = hdr_bytes + sizeof(float) * off[n];
  if (bytes > h->in_cap_b) {
    size_t cap = 1 << 20;
    while (cap < bytes) cap <<= 1;
    cudaFreeHost(h->h_in_b); cudaFree(h->d_in_b);
    h->h_in_b = nullptr; h->d_in_b = nullptr; h->in_cap_b = 0;
    PP_CUDA(cudaMallocHost(&h->h_in_b, cap));
    PP_CUDA(cudaMalloc(&h->d_in_b, cap));
    h->in_cap_b = cap;
  }
  // pack: n headers, then every scene's floats; scenes are independent, so split them over a few threads
  float* payload = reinterpret_cast<float*>(h->h_in_b + hdr_bytes);
  auto pack_range = [&](int k0, int k1) {
    for (int k = k0; k < k1; ++k)
      pack_scene(h, &in[k], reinterpret_cast<PcDevHeader*>(h->h_in_b + kHdrBytes * static_cast<size_t>(k)), payload,
                 static_cast<int>(off[k]));
  };
  const int n_thr = n >= 64 ? std::min<int>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
  if (n_thr == 1) {
    pack_range(0, n);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thr; ++t) pool.emplace_back(pack_range, static_cast<int>(static_cast<long long>(n) * t / n_thr),
                                                      static_cast<int>(static_cast<long long>(n) * (t + 1) / n_thr));
    for (auto& t : pool) t.join();
  }
  const PcScenes S{h->d_in_b, reinterpret_cast<const float*>(h->d_in_b + hdr_bytes), h->d_work_b, h->d_out_b, kHdrBytes};
  PP_CUDA(cudaEventRecord(h->ev0, h->stream));
  PP_CUDA(cudaMemcpyAsync(h->d_in_b, h->h_in_b, bytes, cudaMemcpyHostToDevice, h->stream));
  launch_pass(h, h->stream, S, n);
  pc_front_regions_kernel<<<static_cast<unsigned>(n), 256, 0, h->stream>>>(h->d_out_b, len, static_cast<int>(h->P.width_cells),
                                                                           h->d_regions_b);
  h->launches += h->d_global ? 6 : 5;
  PP_CUDA(cudaGetLastError());
  PP_CUDA(cudaEventRecord(h->ev1, h->stream));
  if (out_host) PP_CUDA(cudaMemcpyAsync(out_host, h->d_out_b, len * n, cudaMemcpyDeviceToHost, h->stream));
  if (out_masks) {
    PP_CUDA(cudaMemcpyAsync(out_masks, h->d_regions_b, sizeof(unsigned) * n, cudaMemcpyDeviceToHost, h->stream));
  }
  PP_CUDA(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  h->batch_n = n;
  return PP_OK;
}

int pc_batch_grids_dev(pc_handle* h, const int8_t** out_dev, int32_t* n_scenes) {
  if (!h || !out_dev) { set_last_error("pc_batch_grids_dev: null argument"); return PP_ERR_INVALID; }
  *out_dev = h->batch_n > 0 ? h->d_out_b : nullptr;
  if (n_scenes) *n_scenes = h->batch_n;
  return PP_OK;
}

int pc_grid_dev(pc_handle* h, const int8_t** out_dev, int32_t* width, int32_t* height) {
  if (!h || !out_dev) { set_last_error("pc_grid_dev: null argument"); return PP_ERR_INVALID; }
  *out_dev = h->d_out;
  if (width) *width = static_cast<int32_t>(h->P.width_cells);
  if (height) *height = static_cast<int32_t>(h->P.height_cells);
  return PP_OK;
}

int pc_front_regions_dev(pc_handle* h, const int8_t* grid_dev, int32_t info_width, int64_t len, uint32_t* out_mask) {
  if (!h || !grid_dev || !out_mask) { set_last_error("pc_front_regions: null argument"); return PP_ERR_INVALID; }
  // the navigator reads info.width * info.width entries (local_navigator.py:82-83, 104-110)
  PP_REQUIRE(info_width >= 0 && info_width <= 16384 && static_cast<int64_t>(info_width) * info_width <= len,
             "pc_front_regions: fewer than width * width cells (the reference runs off msg.data)");
  PP_CUDA(cudaSetDevice(h->device));
  pc_front_regions_kernel<<<1, 256, 0, h->stream>>>(grid_dev, 0, info_width, h->d_regions);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  unsigned m = 0;
  PP_CUDA(cudaMemcpyAsync(&m, h->d_regions, sizeof(m), cudaMemcpyDeviceToHost, h->stream));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  *out_mask = m;
  return PP_OK;
}

int pc_front_regions(pc_handle* h, uint32_t* out_mask) {
  if (!h) { set_last_error("pc_front_regions: null handle"); return PP_ERR_INVALID; }
  return pc_front_regions_dev(h, h->d_out, static_cast<int32_t>(h->P.width_cells), h->G.len, out_mask);
}

int pc_stream(pc_handle* h, void** out_stream) {
  if (!h || !out_stream) return PP_ERR_INVALID;
  *out_stream = h->stream;
  return PP_OK;
}
int pc_launch_count(const pc_handle* h, uint64_t* out) {
  if (!h || !out) return PP_ERR_INVALID;
  *out = h->launches;
  return PP_OK;
}
int pc_last_build_ms(pc_handle* h, float* out_ms) {
  if (!h || !out_ms) return PP_ERR_INVALID;
  *out_ms = h->last_ms;
  return PP_OK;
}

}  // extern "C"
