// pg_global_costmap.cu -- the static global road map on the device (include/prius_costmap.h, row N3).
//
// GlobalCostmap::setupRoadMap of catkin_ws/src/global_costmap/src/global_costmap.cpp ("GC"):
//   readEdges' shift GC:42-43 . constructRoadmapMatrix GC:193-205 . clearMap GC:74-81 .
//   generateRoads GC:91-162 . interpolatePoints GC:164-184 . calcRoadmapMatrixCoords GC:186-191 .
//   calcRoadmap GC:207-217 . calcGridLocation GC:83-89
//
// The reference walks every polyline segment, every interpolated centre-line point, every lane and
// both lane halves sequentially, and lowers a 12000 x 12000 int matrix with `if (m > v) m = v`.
// Since v >= 0 that is m = min(m, trunc(v)) -- order-independent -- so each (centre-line point,
// lane, half) "stroke" becomes one thread doing atomicMin; only the short walk along a stroke stays
// serial, because the cost of a sample depends on how many earlier samples of the SAME stroke were
// inside the matrix (`count++` is skipped for clipped samples, GC:132-140).
//
// The reference's `point` is pair<float, float>: every stored point is rounded to float and
// differences of stored coordinates are float subtractions; everything else is double.  Reproduced
// operation for operation (individually rounded helpers of pp_math.cuh, deterministic
// sin/cos/atan2, -fmad=false).
//
// Kernels: pg_segment_kernel (per polyline segment: direction, length, sample count) -> host scan
// of 10 k counts -> pg_stroke_kernel (5.5 M strokes, ~1.2e8 atomicMin into a 576 MB int32 matrix
// in HBM) -> pg_roadmap_kernel (matrix -> int8 message order: a 32 x 32 shared-memory transpose so
// that both the 576 MB read and the 144 MB write are coalesced; HBM-bound).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/prius_costmap.h"
#include "pp_common.cuh"
#include "pp_math.cuh"

namespace pp {

struct PgSeg {
  float sx, sy;        // shifted start point (GC:42-43)
  double ca, sa;       // cos / sin of angle_bet (GC:169)
  double dist;         // GC:175 with the road extension of GC:170-174
  double cp, sp;       // cos / sin of perp_angle (GC:103)
  int n;               // num_pts_bet (GC:176)
};

__global__ void pg_segment_kernel(const float* __restrict__ xy, const int* __restrict__ seg_first, int n_seg,
                                  pg_params P, PgSeg* __restrict__ segs, int* __restrict__ counts) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const int k = seg_first[s];
  const float sx = __fadd_rn(xy[2 * k], 5.0f), sy = __fsub_rn(xy[2 * k + 1], 5.0f);        // GC:42-43
  const float ex = __fadd_rn(xy[2 * k + 2], 5.0f), ey = __fsub_rn(xy[2 * k + 3], 5.0f);
  double d_x = static_cast<double>(__fsub_rn(sx, ex)), d_y = static_cast<double>(__fsub_rn(sy, ey));   // GC:167-168
  const double angle = dm::atan2_(d_y, d_x);
  PgSeg g;
  g.sx = sx; g.sy = sy;
  dm::sincos(angle, &g.sa, &g.ca);
  d_x = dm::add(d_x, dm::mul(P.road_width, g.ca));                                          // GC:172-173
  d_y = dm::add(d_y, dm::mul(P.road_width, g.sa));
  g.dist = sqrt(dm::add(dm::mul(d_x, d_x), dm::mul(d_y, d_y)));
  if (!dm::trunc_to_int(dm::mul(dm::div(g.dist, P.res), 2.0), &g.n)) g.n = 0;              // GC:176
  if (g.n < 0) g.n = 0;
  dm::sincos(dm::add(angle, 1.5707963267948966), &g.sp, &g.cp);                             // GC:102-103 (M_PI / 2)
  segs[s] = g;
  counts[s] = g.n;
}

// x / res.  When res is a power of two (0.25 in the launch file) the quotient equals the product
// with the exactly representable reciprocal, bit for bit, and costs one multiply instead of a division.
struct PgRes {
  double res, inv;
  bool pow2;
};
__device__ __forceinline__ double pg_div_res(const PgRes& r, double x) { return r.pow2 ? dm::mul(x, r.inv) : dm::div(x, r.res); }

// one stroke (GC:124-157): samples from the lane centre `c` towards the lane edge `e`
__device__ __forceinline__ void pg_stroke(const pg_params& P, const PgRes& R, int H, int W, float cx, float cy, float ex,
                                          float ey, int* __restrict__ matrix) {
  const double d_x = static_cast<double>(__fsub_rn(cx, ex)), d_y = static_cast<double>(__fsub_rn(cy, ey));
  const double angle = dm::atan2_(d_y, d_x);
  double sa, ca;
  dm::sincos(angle, &sa, &ca);
  const double dist = sqrt(dm::add(dm::mul(d_x, d_x), dm::mul(d_y, d_y)));
  int n;
  if (!dm::trunc_to_int(dm::mul(dm::div(dist, P.res), 2.0), &n)) n = 0;
  int count = 0;
  const double nd = static_cast<double>(n);
  const double half_h = dm::div(P.height_cells, 2.0), half_w = dm::div(P.width_cells, 2.0);
  for (int i = 0; i < n; ++i) {
    const double qi = dm::div(static_cast<double>(i), nd);
    const double t = dm::mul(qi, dist);
    const float px = static_cast<float>(dm::add(static_cast<double>(cx), dm::mul(t, ca)));   // GC:179-181, stored as float
    const float py = static_cast<float>(dm::add(static_cast<double>(cy), dm::mul(t, sa)));
    // count == i unless an earlier sample of this stroke was clipped at the map border
    const double qc = count == i ? qi : dm::div(static_cast<double>(count), nd);
    const double value = dm::sub(99.0, dm::mul(qc, 99.0));                                   // GC:130
    int xi, yi;
    if (!dm::trunc_to_int(dm::sub(half_h, pg_div_res(R, static_cast<double>(px))), &xi)) continue;   // GC:188-189
    if (!dm::trunc_to_int(dm::sub(half_w, pg_div_res(R, static_cast<double>(py))), &yi)) continue;
    if (xi < 0 || xi > H - 1 || yi < 0 || yi > W - 1) continue;                              // GC:132-135
    atomicMin(&matrix[static_cast<size_t>(xi) * W + yi], __double2int_rz(value));             // GC:136-139
    count++;
  }
}

__global__ void __launch_bounds__(256)
pg_stroke_kernel(const PgSeg* __restrict__ segs, const long long* __restrict__ seg_off, int n_seg, long long n_points,
                 pg_params P, PgRes R, int H, int W, int* __restrict__ matrix) {
  const long long strokes = n_points * P.num_lanes * 2;
  const double lane_width = dm::div(P.road_width, static_cast<double>(P.num_lanes));          // GC:104
  const double first_off = dm::add(dm::div(P.road_width, 2.0), dm::div(lane_width, 2.0));     // GC:108
  const double half_lane = dm::div(lane_width, 2.0);
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < strokes;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long point = t / (P.num_lanes * 2);
    const int rem = static_cast<int>(t % (P.num_lanes * 2));
    const int lane = rem >> 1, side = rem & 1;
    // segment of this centre-line point: last s with seg_off[s] <= point
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (seg_off[mid] <= point) lo = mid; else hi = mid - 1;
    }
    const PgSeg g = segs[lo];
    const int i = static_cast<int>(point - seg_off[lo]);
    const double tt = dm::mul(dm::div(static_cast<double>(i), static_cast<double>(g.n)), g.dist);
    const float px = static_cast<float>(dm::add(static_cast<double>(g.sx), dm::mul(tt, g.ca)));   // GC:179-181
    const float py = static_cast<float>(dm::add(static_cast<double>(g.sy), dm::mul(tt, g.sa)));
    const double first_x = dm::sub(static_cast<double>(px), dm::mul(first_off, g.cp));            // GC:108-109
    const double first_y = dm::sub(static_cast<double>(py), dm::mul(first_off, g.sp));
    const double jl = dm::mul(static_cast<double>(lane), lane_width);
    const float lcx = static_cast<float>(dm::add(first_x, dm::mul(jl, g.cp)));                    // GC:112-114
    const float lcy = static_cast<float>(dm::add(first_y, dm::mul(jl, g.sp)));
    float ex, ey;
    if (side == 0) {                                                                              // GC:118-119
      ex = static_cast<float>(dm::sub(static_cast<double>(lcx), dm::mul(half_lane, g.cp)));
      ey = static_cast<float>(dm::sub(static_cast<double>(lcy), dm::mul(half_lane, g.sp)));
    } else {                                                                                      // GC:120-121
      ex = static_cast<float>(dm::add(static_cast<double>(lcx), dm::mul(half_lane, g.cp)));
      ey = static_cast<float>(dm::add(static_cast<double>(lcy), dm::mul(half_lane, g.sp)));
    }
    pg_stroke(P, R, H, W, lcx, lcy, ex, ey, matrix);
  }
}

// calcRoadmap (GC:207-217): out[location(x, y)] = matrix[x][y].  A 64 x 64 tile goes through shared
// memory: 16-byte loads along y (the matrix's fast axis, four of them in flight per thread), byte
// stores along x (the message's fast axis: `location` falls by one when x grows by one).
__global__ void __launch_bounds__(256)
pg_roadmap_kernel(const int* __restrict__ matrix, int H, int W, pg_params P, int8_t* __restrict__ out, long long out_len) {
  __shared__ int tile[64][65];
  const int x0 = blockIdx.y * 64, y0 = blockIdx.x * 64;
  const int t = threadIdx.x;
  const bool vec = (W & 3) == 0;                               // rows stay 16-byte aligned
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {                       // 16 rows x 16 int4 per pass
    const int r = pass * 16 + (t >> 4), c = (t & 15) * 4;
    const int x = x0 + r, y = y0 + c;
    int4 v = make_int4(99, 99, 99, 99);
    if (x < H) {
      const int* src = matrix + static_cast<size_t>(x) * W + y;
      if (vec && y + 3 < W) v = *reinterpret_cast<const int4*>(src);
      else {
        if (y < W) v.x = src[0];
        if (y + 1 < W) v.y = src[1];
        if (y + 2 < W) v.z = src[2];
        if (y + 3 < W) v.w = src[3];
      }
    }
    tile[r][c] = v.x; tile[r][c + 1] = v.y; tile[r][c + 2] = v.z; tile[r][c + 3] = v.w;
  }
  __syncthreads();
#pragma unroll 4
  for (int pass = 0; pass < 16; ++pass) {                      // 4 columns x 64 consecutive x per pass
    const int r = pass * 4 + (t >> 6), xl = t & 63;
    const int y = y0 + r, x = x0 + xl;
    if (x >= H || y >= W) continue;
    int wp, hp, loc;
    if (!dm::trunc_to_int(dm::sub(P.width_cells, static_cast<double>(static_cast<float>(y))), &wp)) continue;    // GC:85
    if (!dm::trunc_to_int(dm::sub(P.height_cells, static_cast<double>(static_cast<float>(x))), &hp)) continue;   // GC:86
    if (!dm::trunc_to_int(dm::add(static_cast<double>(hp), dm::mul(static_cast<double>(wp), P.height_cells)), &loc)) continue;
    if (loc < 0 || loc >= out_len) continue;                   // the reference writes past the end here
    out[loc] = static_cast<int8_t>(tile[xl][r]);
  }
}

__global__ void pg_fill_kernel(int4* __restrict__ p, size_t n4, int v) {
  const int4 q = make_int4(v, v, v, v);
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n4; k += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[k] = q;
}

static thread_local float g_pg_ms = 0;
static thread_local int g_pg_launches = 0;

}  // namespace pp

using namespace pp;

extern "C" {

int pg_default_params(pg_params* p) {
  if (!p) return PP_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->res = 0.25;                 // full_sim.launch:77
  p->road_width = 11.1;          // :69
  p->num_lanes = 2;              // :70
  p->height_cells = 3000 / 0.25; // :71, GC:14
  p->width_cells = 3000 / 0.25;  // :72, GC:15
  return PP_OK;
}

int pg_last_build(float* out_ms, int32_t* out_launches) {
  if (out_ms) *out_ms = g_pg_ms;
  if (out_launches) *out_launches = g_pg_launches;
  return PP_OK;
}

int pg_free(int device, int8_t* grid_dev) {
  if (!grid_dev) return PP_OK;
  PP_CUDA(cudaSetDevice(device));
  PP_CUDA(cudaFree(grid_dev));
  return PP_OK;
}

int pg_build(const pg_params* params, int device, const float* xy_host, const int32_t* offsets_host,
             int32_t n_sections, int8_t* out_host, int8_t** out_dev) {
  if (!params || !out_dev) { set_last_error("pg_build: null argument"); return PP_ERR_INVALID; }
  *out_dev = nullptr;
  PP_REQUIRE(n_sections >= 0 && (n_sections == 0 || (xy_host && offsets_host)), "pg_build: bad polylines");
  PP_REQUIRE(params->res > 0 && params->num_lanes >= 1 && params->num_lanes <= 64 && params->road_width > 0,
             "pg_build: bad parameters");
  PP_REQUIRE(params->height_cells >= 1 && params->width_cells >= 1 &&
             params->height_cells * params->width_cells < 1073741824.0, "pg_build: bad grid size");
  PP_REQUIRE(std::floor(params->height_cells) == params->height_cells && std::floor(params->width_cells) == params->width_cells,
             "pg_build: cell counts must be whole numbers (metres / res of the launch file are)");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    set_last_error(std::string("pg_build: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    return PP_ERR_CUDA;
  }
  PP_REQUIRE(device >= 0 && device < n_dev, "pg_build: bad device index");
  PP_CUDA(cudaSetDevice(device));
  const pg_params P = *params;
  const int H = static_cast<int>(P.height_cells), W = static_cast<int>(P.width_cells);     // GC:195-196
  long long out_len = 0;                                                                   // GC:77: i < H * W
  { const double prod = P.height_cells * P.width_cells; out_len = static_cast<long long>(std::ceil(prod)); }
  const int n_pts = n_sections ? offsets_host[n_sections] : 0;
  std::vector<int> seg_first;
  for (int s = 0; s < n_sections; ++s) {
    PP_REQUIRE(offsets_host[s + 1] >= offsets_host[s], "pg_build: offsets must be non-decreasing");
    for (int k = offsets_host[s]; k + 1 < offsets_host[s + 1]; ++k) seg_first.push_back(k);   // GC:95-98
  }
  const int n_seg = static_cast<int>(seg_first.size());
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float* d_xy = nullptr; int* d_first = nullptr; PgSeg* d_seg = nullptr; int* d_cnt = nullptr;
  long long* d_off = nullptr; int* d_matrix = nullptr; int8_t* d_out = nullptr;
  int rc = PP_OK, launches = 0;
  auto cleanup = [&]() {
    cudaFree(d_xy); cudaFree(d_first); cudaFree(d_seg); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_matrix);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
  };
#define PG_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); cudaFree(d_out); cleanup(); return PP_ERR_CUDA; } } while (0)
  PG_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  PG_TRY(cudaEventCreate(&e0));
  PG_TRY(cudaEventCreate(&e1));
  const size_t cells = static_cast<size_t>(H) * W;
  const size_t cells4 = (cells + 3) / 4 * 4;
  PG_TRY(cudaMalloc(&d_matrix, sizeof(int) * cells4));
  PG_TRY(cudaMalloc(&d_out, static_cast<size_t>(out_len) + 16));
  PG_TRY(cudaEventRecord(e0, st));
  pg_fill_kernel<<<sm * 16, 256, 0, st>>>(reinterpret_cast<int4*>(d_matrix), cells4 / 4, 99);    // GC:193-205
  launches++;
  PG_TRY(cudaMemsetAsync(d_out, 0xFF, static_cast<size_t>(out_len), st));                        // GC:74-81: -1
  std::vector<long long> off(n_seg + 1, 0);
  if (n_seg > 0) {
    PG_TRY(cudaMalloc(&d_xy, sizeof(float) * 2 * static_cast<size_t>(n_pts)));
    PG_TRY(cudaMalloc(&d_first, sizeof(int) * n_seg));
    PG_TRY(cudaMalloc(&d_seg, sizeof(PgSeg) * n_seg));
    PG_TRY(cudaMalloc(&d_cnt, sizeof(int) * n_seg));
    PG_TRY(cudaMalloc(&d_off, sizeof(long long) * (n_seg + 1)));
    PG_TRY(cudaMemcpyAsync(d_xy, xy_host, sizeof(float) * 2 * static_cast<size_t>(n_pts), cudaMemcpyHostToDevice, st));
    PG_TRY(cudaMemcpyAsync(d_first, seg_first.data(), sizeof(int) * n_seg, cudaMemcpyHostToDevice, st));
    pg_segment_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(d_xy, d_first, n_seg, P, d_seg, d_cnt);
    launches++;
    std::vector<int> cnt(n_seg);
    PG_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * n_seg, cudaMemcpyDeviceToHost, st));
    PG_TRY(cudaStreamSynchronize(st));
    for (int s = 0; s < n_seg; ++s) off[s + 1] = off[s] + cnt[s];
    PG_TRY(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * (n_seg + 1), cudaMemcpyHostToDevice, st));
    const long long strokes = off[n_seg] * P.num_lanes * 2;
    if (strokes > 0) {
      const int blocks = static_cast<int>(std::min<long long>((strokes + 255) / 256, static_cast<long long>(sm) * 64));
      PgRes R;
      int ex2 = 0;
      R.res = P.res;
      R.pow2 = std::frexp(P.res, &ex2) == 0.5 && ex2 > -900 && ex2 < 900;
      R.inv = R.pow2 ? 1.0 / P.res : 0.0;
      pg_stroke_kernel<<<blocks, 256, 0, st>>>(d_seg, d_off, n_seg, off[n_seg], P, R, H, W, d_matrix);
      launches++;
    }
  }
  {
    dim3 grid((W + 63) / 64, (H + 63) / 64);
    pg_roadmap_kernel<<<grid, 256, 0, st>>>(d_matrix, H, W, P, d_out, out_len);
    launches++;
  }
  PG_TRY(cudaGetLastError());
  PG_TRY(cudaEventRecord(e1, st));
  if (out_host) PG_TRY(cudaMemcpyAsync(out_host, d_out, static_cast<size_t>(out_len), cudaMemcpyDeviceToHost, st));
  PG_TRY(cudaStreamSynchronize(st));
#undef PG_TRY
  cudaEventElapsedTime(&g_pg_ms, e0, e1);
  g_pg_launches = launches;
  cleanup();
  *out_dev = d_out;
  return rc;
}

}  // extern "C"
