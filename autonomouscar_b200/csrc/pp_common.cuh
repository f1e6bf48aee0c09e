// pp_common.cuh -- handle layout, device-side grid representations and the per-pose
// collision primitives shared by the batch kernels and the planner kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda_runtime.h>

#include "../../include/prius_planner.h"
#include "pp_math.cuh"

namespace pp {

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);
#define PP_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ::pp::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(e_));           \
      return PP_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)
#define PP_REQUIRE(cond, msg)                                                             \
  do {                                                                                    \
    if (!(cond)) { ::pp::set_last_error(msg); return PP_ERR_INVALID; }                    \
  } while (0)

// ---------------------------------------------------------------------------------------
// footprint sample lattice (row S3), computed once on the host at pp_create
// ---------------------------------------------------------------------------------------
struct Lattice {
  int nu, nv;      // sample points along the car length / across the width
  double su, sv;   // lattice step, in cells
  double hu, hv;   // (nu-1)/2, (nv-1)/2
};
constexpr int kAnchorOffset = 48;         // anchor cell = floor(centre) - 48
constexpr int kFracBits = 24;             // Q7.24 cell coordinates relative to the anchor
constexpr double kFixedOne = 16777216.0;

// ---------------------------------------------------------------------------------------
// device-side grid representations, built by pp_set_grid (pp_grid.cu)
// ---------------------------------------------------------------------------------------
#ifndef PP_NEAR_BINS
#define PP_NEAR_BINS 16
#endif
constexpr int kNearBins = PP_NEAR_BINS;               // yaw bins of the batch kernel's first test (GridDev::near4)
constexpr int kHaloCells = 128;           // zero halo around the occupancy tiles
constexpr int kTile = 32;                 // a tile is 32 x 32 cells = 32 words = one 128 B line

struct GridDev {
  int W, H;                 // m_c_space dimensions: first index i < W, second index j < H
  double res;
  int8_t* cspace;           // W*H bytes, cspace[i*H + j] == m_c_space[i][j]  (LP:425-441)
  // occupancy (cell == 100) as 32x32-bit tiles over cells [-halo, W+halo) x [-halo, H+halo):
  //   word r of tile (ti,tj) = row i = 32*ti - halo + r, bit b = column j = 32*tj - halo + b
  uint32_t* occ_tiles;
  uint8_t* tile_any;        // 1 if any bit of the tile is set
  int TI, TJ;               // tile counts
  // one bit per 8x8-cell block of the haloed domain ("any cell of the block occupied"):
  //   coarse[ci8 * CW8 + (cj8 >> 6)] bit (cj8 & 63), ci8 = (i + halo) >> 3, cj8 = (j + halo) >> 3
  unsigned long long* coarse;
  int CI8, CW8;
  // the same per 4x4-cell block (tighter pre-test of the batch kernel's bounding box): ci4 = (i + halo) >> 2
  unsigned long long* coarse4;   // 4x4-cell blocks, word pairs: {any cell occupied, all 16 cells occupied}
  int CI4, CW4;
  // kNearBins maps, one per yaw bin of pi / kNearBins (the footprint is symmetric about the pose, so yaw and
  // yaw + pi share a bin):  near4[(bin * CI4 + ci4) * CW4 + (cj4 >> 6)] bit (cj4 & 63) is set <=> some obstacle
  // lies in a 4x4 block that the footprint can touch, for a pose anywhere in block (ci4, cj4) at any yaw of the
  // bin -- the "any" map dilated by the bin's structuring element (pp_grid.cu: near_span_table).  The batch
  // kernel's first test: one bit per pose.
  unsigned long long* near4;
  int reach_cells;          // footprint half diagonal + slack, in cells: poses further than this from the grid are free
  // REF_AABB box-dilated map over base cells [-aabb_pad_x, W+aabb_pad_x) x [-aabb_pad_y, H+...)
  //   aabb_map[(bx+pad_x) * aabb_stride + (by+pad_y)] = 1 iff the LP:477-495 box anchored at
  //   (bx,by) contains a cell == 100
  uint8_t* aabb_map;
  int aabb_pad_x, aabb_pad_y, aabb_nx, aabb_ny;
  // aabb_map summarised per block of 2^aabb_cshift x 2^aabb_cshift map cells, 2 bits per block
  // (bit 0: some cell of the block is 1, bit 1: the block lies inside the map and all its cells are 1),
  // 16 blocks per word along my: word (mx >> s) * aabb_cwpr + (my >> s >> 4).  Small enough for shared
  // memory (<= 24 KB), so the batch kernel answers ~90 % of the poses without touching L2.
  uint32_t* aabb_coarse;
  int aabb_cshift, aabb_cnbx, aabb_cwpr;
};

// A bank = n grids of ONE shape stored back to back: every array of grid k starts k * stride elements
// after grid 0's.  The planner's own grid is a bank of one (strides unused).  pp_set_grid_bank_dev builds
// banks; pp_plan_batch_banked lets every query name its grid.
struct GridBank {
  int n = 0;
  size_t cspace = 0, occ_tiles = 0, tile_any = 0, coarse = 0, coarse4 = 0, near4 = 0, aabb_map = 0, aabb_coarse = 0;
};
__host__ __device__ inline GridDev grid_at(GridDev g, const GridBank& b, unsigned k) {
  g.cspace += k * b.cspace;
  g.occ_tiles += k * b.occ_tiles;
  g.tile_any += k * b.tile_any;
  g.coarse += k * b.coarse;
  g.coarse4 += k * b.coarse4;
  g.near4 += k * b.near4;
  g.aabb_map += k * b.aabb_map;
  g.aabb_coarse += k * b.aabb_coarse;
  return g;
}

struct CollideParams {
  Lattice lat;
  double car_length, car_width, search_res;
  int mode;
};

// ---------------------------------------------------------------------------------------
// per-pose primitives
// ---------------------------------------------------------------------------------------
struct PoseFx {
  int valid;
  int ai, aj;            // anchor cell
  int32_t oi, oj;        // lattice origin (a=0,b=0) relative to the anchor, Q7.24
  int32_t ui, uj;        // step along the length
  int32_t vi, vj;        // step across the width
};

__device__ __forceinline__ int32_t to_fixed(double v) {
  return __double2int_rn(dm::mul(v, kFixedOne));
}

// Mirrors oracle/collision_ref.h pose_to_fixed operation for operation.
__device__ __forceinline__ PoseFx pose_to_fixed(const Lattice& l, int W, int H, double res, double x,
                                                double y, double yaw) {
  PoseFx p;
  p.valid = 0;
  double s, c;
  dm::sincos(yaw, &s, &c);
  const double ci = dm::sub(static_cast<double>(W / 2), dm::div(y, res));
  const double cj = dm::sub(static_cast<double>(H / 2), dm::div(x, res));
  if (!(fabs(ci) < 1048576.0) || !(fabs(cj) < 1048576.0) || s != s) return p;
  p.ai = __double2int_rd(ci) - kAnchorOffset;
  p.aj = __double2int_rd(cj) - kAnchorOffset;
  const double ui = -dm::mul(s, l.su), uj = -dm::mul(c, l.su);
  const double vi = -dm::mul(c, l.sv), vj = dm::mul(s, l.sv);
  const double oi = dm::sub(dm::sub(dm::sub(ci, static_cast<double>(p.ai)), dm::mul(l.hu, ui)), dm::mul(l.hv, vi));
  const double oj = dm::sub(dm::sub(dm::sub(cj, static_cast<double>(p.aj)), dm::mul(l.hu, uj)), dm::mul(l.hv, vj));
  p.oi = to_fixed(oi); p.oj = to_fixed(oj);
  p.ui = to_fixed(ui); p.uj = to_fixed(uj);
  p.vi = to_fixed(vi); p.vj = to_fixed(vj);
  p.valid = 1;
  return p;
}

// cell range touched by the lattice (min/max over its four corners), in grid cells
__device__ __forceinline__ void lattice_bbox(const Lattice& l, const PoseFx& p, int* i0, int* i1, int* j0,
                                             int* j1) {
  const int32_t eu_i = (l.nu - 1) * p.ui, eu_j = (l.nu - 1) * p.uj;
  const int32_t ev_i = (l.nv - 1) * p.vi, ev_j = (l.nv - 1) * p.vj;
  const int32_t ci0 = p.oi, ci1 = p.oi + eu_i, ci2 = p.oi + ev_i, ci3 = p.oi + eu_i + ev_i;
  const int32_t cj0 = p.oj, cj1 = p.oj + eu_j, cj2 = p.oj + ev_j, cj3 = p.oj + eu_j + ev_j;
  *i0 = p.ai + (min(min(ci0, ci1), min(ci2, ci3)) >> kFracBits);
  *i1 = p.ai + (max(max(ci0, ci1), max(ci2, ci3)) >> kFracBits);
  *j0 = p.aj + (min(min(cj0, cj1), min(cj2, cj3)) >> kFracBits);
  *j1 = p.aj + (max(max(cj0, cj1), max(cj2, cj3)) >> kFracBits);
}

__device__ __forceinline__ uint32_t occ_word(const GridDev& g, int i, int jw_tile) {
  // i: cell row (may be in the halo), jw_tile: tile column index
  const int ih = i + kHaloCells;
  return __ldg(&g.occ_tiles[(static_cast<size_t>(ih >> 5) * g.TJ + jw_tile) * kTile + (ih & 31)]);
}

// 8x8-cell coarse occupancy under the cell box [i0,i1] x [j0,j1] (<= 8 x 8 coarse cells)
__device__ __forceinline__ bool coarse_box_any(const GridDev& g, int i0, int i1, int j0, int j1) {
  const int r0 = (i0 + kHaloCells) >> 3, r1 = (i1 + kHaloCells) >> 3;
  const int c0 = (j0 + kHaloCells) >> 3, c1 = (j1 + kHaloCells) >> 3;
  const int w0 = c0 >> 6, sh = c0 & 63, nb = c1 - c0 + 1;            // nb <= 8 bits starting at bit c0
  const unsigned long long m = ((1ull << nb) - 1ull);
  unsigned long long any = 0;
  for (int r = r0; r <= r1; ++r) {
    const unsigned long long* row = g.coarse + static_cast<size_t>(r) * g.CW8 + w0;
    unsigned long long bits = __ldg(row) >> sh;
    if (sh + nb > 64) bits |= __ldg(row + 1) << (64 - sh);
    any |= bits & m;
  }
  return any != 0;
}

// The footprint as two BANDS of columns per row.  The samples are O + a u + b v = O + s U + t V with U = (nu-1) u,
// V = (nv-1) v and s, t in [0, 1], so with D = (di, dj) the offset from O (rows, columns) and X = cross(U, V):
//     t X = Ui dj - Uj di   and   s X = di Vj - dj Vi,   i.e.
//     dj in kT di + [eT0, eT1]  (kT = Uj/Ui, {0, X/Ui})     and     dj in kS di + [eS0, eS1]  (kS = Vj/Vi, {0, -X/Vi})
// -- each band linear in di.  Over a strip of rows a band's HULL (columns it reaches for some di of the strip) and
// its CORE (columns inside it for every di of the strip) are its values at one end of the strip or the other.  A
// band whose divisor is under half a cell (an edge nearly parallel to the columns) is dropped: it is unbounded in
// the hull (the bounding box still applies) and no core is formed.  Users widen hulls and narrow cores by
// kBandMargin = 1/8 cell; all values are relative to the box or the window (< 1e4 cells in magnitude, slopes
// <= 2 * 48 / 0.5), so float rounding stays under 0.01 cell and what the bands prove stays proven.
#ifndef PP_COARSE4_UNROLL
#define PP_COARSE4_UNROLL 1
#endif
constexpr int kCoarse4Unroll = PP_COARSE4_UNROLL;
struct Bands {
  float kT, eT0, eT1, kS, eS0, eS1;
  int core;          // both bands present and the lattice dense enough for the full-block argument below
};
constexpr float kBandMargin = 0.125f, kBandBig = 1.0e9f;

__device__ __forceinline__ Bands footprint_bands(const Lattice& l, const PoseFx& p) {
  constexpr float q = 1.0f / static_cast<float>(1 << kFracBits);
  const float ui = static_cast<float>(p.ui) * q, uj = static_cast<float>(p.uj) * q;
  const float vi = static_cast<float>(p.vi) * q, vj = static_cast<float>(p.vj) * q;
  const float Ui = static_cast<float>(l.nu - 1) * ui, Uj = static_cast<float>(l.nu - 1) * uj;
  const float Vi = static_cast<float>(l.nv - 1) * vi, Vj = static_cast<float>(l.nv - 1) * vj;
  const float X = Ui * Vj - Uj * Vi;
  Bands b;
  b.kT = 0.0f; b.eT0 = -kBandBig; b.eT1 = kBandBig;
  b.kS = 0.0f; b.eS0 = -kBandBig; b.eS1 = kBandBig;
  const bool hasT = fabsf(Ui) >= 0.5f, hasS = fabsf(Vi) >= 0.5f;
  if (hasT) {
    const float inv = __fdividef(1.0f, Ui), e = X * inv;
    b.kT = Uj * inv; b.eT0 = fminf(0.0f, e); b.eT1 = fmaxf(0.0f, e);
  }
  if (hasS) {
    const float inv = __fdividef(1.0f, Vi), e = -X * inv;
    b.kS = Vj * inv; b.eS0 = fminf(0.0f, e); b.eS1 = fmaxf(0.0f, e);
  }
  b.core = hasT && hasS && ui * ui + uj * uj + vi * vi + vj * vj <= 4.0f;
  return b;
}

// The 4x4-cell summaries under the ORIENTED footprint instead of its bounding box.  Returns 0 = certainly free,
// 2 = certainly colliding, 1 = the narrow phase has to look.  Per coarse row (a strip of 4 cell rows, widened by
// the margin on either side):
//   * the intersection of the two hulls, widened by the margin, holds every sample of the row:  no "any" bit
//     there in any row -> free (the bounding box is 2.6 times the footprint's area for a car at 45 degrees);
//   * a 4x4 block whose columns lie in both cores, narrowed by the margin, lies inside the footprint; if the block
//     is fully occupied ("full" bit) a sample lands in it -- the lattice's covering radius sqrt(|u|^2 + |v|^2) / 2
//     is below the block's inscribed radius of 2 cells whenever |u|^2 + |v|^2 <= 4 (Bands::core), and the lattice
//     point nearest the block's centre is then inside the block, hence inside the footprint, hence one of the
//     samples -> collision.
// Both verdicts are proofs; every undecided pose takes the exact test.
// *tile_rows gets bit a set when tile row (first tile row under the box) + a holds an obstacle inside the box's
// columns: the narrow phase loads only those rows of its window.
// kUnroll: 1 where many warps hide the latency of a strip's load (the batch kernel: the trip counts differ per lane
// and unrolled copies cost more than they hide), 4 where a few warps wait for one pose (edges of the sampling planner).
template <int kUnroll>
__device__ __forceinline__ int coarse4_oriented_test(const GridDev& g, const Bands& B, const PoseFx& p, int i0,
                                                     int i1, int j0, int j1, uint32_t* tile_rows) {
  const int r0 = (i0 + kHaloCells) >> 2, r1 = (i1 + kHaloCells) >> 2;
  const int c0 = (j0 + kHaloCells) >> 2, c1 = (j1 + kHaloCells) >> 2;
  const int w0 = c0 >> 6, sh = c0 & 63, nb = c1 - c0 + 1;            // nb <= 16 bits starting at bit c0
  constexpr float q = 1.0f / static_cast<float>(1 << kFracBits);
  constexpr float kM = kBandMargin, kStrip = 4.0f + 2.0f * kM;
  // columns in cells from the first coarse column of the box
  const float jbase = static_cast<float>(p.aj + kHaloCells - 4 * c0) + static_cast<float>(p.oj) * q;
  const float wT = kStrip * B.kT, wS = kStrip * B.kS;
  // hull [h0, h1] and core [k0, k1] of each band as offsets to slope * d
  const float th0 = B.eT0 + fminf(0.0f, wT) + (jbase - kM), th1 = B.eT1 + fmaxf(0.0f, wT) + (jbase + kM);
  const float sh0 = B.eS0 + fminf(0.0f, wS) + (jbase - kM), sh1 = B.eS1 + fmaxf(0.0f, wS) + (jbase + kM);
  const float tk0 = B.core ? B.eT0 + fmaxf(0.0f, wT) + (jbase + kM) : kBandBig;    // (no core: empty in every row)
  const float tk1 = B.eT1 + fminf(0.0f, wT) + (jbase - kM);
  const float sk0 = B.core ? B.eS0 + fmaxf(0.0f, wS) + (jbase + kM) : kBandBig;
  const float sk1 = B.eS1 + fminf(0.0f, wS) + (jbase - kM);
  const float dbase = kM + static_cast<float>(p.oi) * q;             // di of the strip's first row = 4 r - halo - ai - dbase
  const float jhi = static_cast<float>(4 * nb - 1);
  uint32_t any = 0, full = 0, rows = 0;
  const uint32_t box = (2u << (nb - 1)) - 1u;                        // all columns of the box
  const int tr0 = r0 >> 3;                                           // (8 coarse rows per 32-cell tile)
  const bool two = sh + nb > 64;                                     // the box's columns straddle two words
  // straight-line body (no data-dependent branch): the loads of successive rows overlap.  One 128-bit load brings
  // the "any" and the "full" word of a row (they are interleaved in memory).
  const ulonglong2* map = reinterpret_cast<const ulonglong2*>(g.coarse4);
#pragma unroll kUnroll
  for (int r = r0; r <= r1; ++r) {
    const float d = static_cast<float>(4 * r - kHaloCells - p.ai) - dbase;
    const float td = B.kT * d, sd = B.kS * d;
    const float lo = fmaxf(fmaxf(td + th0, sd + sh0), 0.0f), hi = fminf(fminf(td + th1, sd + sh1), jhi);
    // block c (cells 4c .. 4c+3 from the box's first coarse column) is inside iff 4c >= klo and 4c + 4 <= khi
    const float klo = fmaxf(fmaxf(td + tk0, sd + sk0), -4.0f), khi = fminf(fminf(td + tk1, sd + sk1), jhi + 1.0f);
    const int a = static_cast<int>(lo) >> 2, b = static_cast<int>(hi) >> 2;   // 0 <= a <= b < nb when lo <= hi
    const int ca = __float2int_rd(klo * 0.25f) + 1, cb = __float2int_rd(khi * 0.25f) - 1;
    const uint32_t hm = (lo <= hi) ? (((2u << b) - 1u) & (0xffffffffu << a)) : 0u;
    const uint32_t km = (klo <= khi && ca <= cb) ? (((2u << cb) - 1u) & (0xffffffffu << ca)) : 0u;
    const ulonglong2* row = map + static_cast<size_t>(r) * g.CW4 + w0;
    const ulonglong2 w = __ldg(row);
    unsigned long long bits = w.x >> sh, fb = w.y >> sh;
    if (two) {
      const ulonglong2 w2 = __ldg(row + 1);
      bits |= w2.x << (64 - sh);
      fb |= w2.y << (64 - sh);
    }
    any |= static_cast<uint32_t>(bits) & hm;
    full |= static_cast<uint32_t>(fb) & km;
    rows |= (static_cast<uint32_t>(bits) & box) ? 1u << ((r >> 3) - tr0) : 0u;
  }
  *tile_rows = rows;
  return full ? 2 : (any ? 1 : 0);
}

// ORIENTED, one thread evaluating the whole lattice straight from the tiled bitmap.
// Used where poses are few and latency-bound (planner candidates, edge steps); the batch
// path uses the warp-cooperative kernel in pp_collide.cu.
__device__ __forceinline__ bool oriented_hit_thread(const GridDev& g, const Lattice& l, const PoseFx& p) {
  if (!p.valid) return false;
  int i0, i1, j0, j1;
  lattice_bbox(l, p, &i0, &i1, &j0, &j1);
  if (i1 < 0 || i0 >= g.W || j1 < 0 || j0 >= g.H) return false;
  // exact conservative pruning first (same tests as the batch kernel's broad phase): no obstacle
  // in the tiles / 8x8 blocks under the bounding box means no sample can hit one
  {
    const int t0i = (i0 + kHaloCells) >> 5, t1i = (i1 + kHaloCells) >> 5;
    const int t0j = (j0 + kHaloCells) >> 5, t1j = (j1 + kHaloCells) >> 5;
    uint32_t any = 0;
    for (int a = t0i; a <= t1i; ++a)
      for (int b = t0j; b <= t1j; ++b) any |= __ldg(&g.tile_any[static_cast<size_t>(a) * g.TJ + b]);
    if (!any) return false;
    if (!coarse_box_any(g, i0, i1, j0, j1)) return false;
  }
  // all loads of four lattice columns are independent (no early exit in between), so they
  // overlap in the memory pipeline; the exit test runs once per four columns
  uint32_t hit = 0;
  const size_t tj_stride = kTile;
  for (int a0 = 0; a0 < l.nu; a0 += 4) {
    const int a1 = min(a0 + 4, l.nu);
    for (int a = a0; a < a1; ++a) {
      const int32_t ci = p.oi + a * p.ui, cj = p.oj + a * p.uj;
#pragma unroll 4
      for (int b = 0; b < l.nv; ++b) {
        const int32_t fi = ci + b * p.vi, fj = cj + b * p.vj;
        const int ih = p.ai + (fi >> kFracBits) + kHaloCells, jh = p.aj + (fj >> kFracBits) + kHaloCells;
        const uint32_t w = __ldg(&g.occ_tiles[(static_cast<size_t>(ih >> 5) * g.TJ + (jh >> 5)) * tj_stride + (ih & 31)]);
        hit |= w >> (jh & 31);
      }
    }
    if (hit & 1u) return true;
  }
  return false;
}

// REF_AABB through the dilated map: LP:471-476 evaluated verbatim for the base cell, then
// one lookup replaces the LP:477-495 / LP:447-456 loops.
__device__ __forceinline__ bool aabb_hit(const GridDev& g, double px, double py) {
  int bx, by;
  if (!dm::trunc_to_int(dm::sub(static_cast<double>(g.W / 2), dm::div(py, g.res)), &bx)) return false;
  if (!dm::trunc_to_int(dm::sub(static_cast<double>(g.H / 2), dm::div(px, g.res)), &by)) return false;
  const int mx = bx + g.aabb_pad_x, my = by + g.aabb_pad_y;
  if (mx < 0 || mx >= g.aabb_nx || my < 0 || my >= g.aabb_ny) return false;
  return __ldg(&g.aabb_map[static_cast<size_t>(mx) * g.aabb_ny + my]) != 0;
}

__device__ __forceinline__ bool collide_pose_thread(const GridDev& g, const CollideParams& cp, double x,
                                                    double y, double yaw) {
  if (cp.mode == PP_COLLISION_AS_SHIPPED) return false;  // LP:445
  if (cp.mode == PP_COLLISION_REF_AABB) return aabb_hit(g, x, y);
  return oriented_hit_thread(g, cp.lat, pose_to_fixed(cp.lat, g.W, g.H, g.res, x, y, yaw));
}

// along-edge rule, mirrors oracle collide_edge
__device__ __forceinline__ bool collide_edge_thread(const GridDev& g, const CollideParams& cp, double x0,
                                                    double y0, double x1, double y1, double yaw) {
  const double dx = dm::sub(x1, x0), dy = dm::sub(y1, y0);
  const double dist = __dsqrt_rn(dm::add(dm::mul(dx, dx), dm::mul(dy, dy)));
  int n;
  if (!dm::trunc_to_int(dm::mul(dm::div(dist, cp.search_res), 2.0), &n)) return false;
  if (n < 1) n = 1;
  for (int k = 1; k <= n; ++k) {
    const double t = dm::div(static_cast<double>(k), static_cast<double>(n));
    const double px = dm::add(x0, dm::mul(t, dx)), py = dm::add(y0, dm::mul(t, dy));
    if (collide_pose_thread(g, cp, px, py, yaw)) return true;
  }
  return false;
}

}  // namespace pp

// ---------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------
struct PlanWorkspace;   // pp_plan.cu
struct RrtWorkspace;    // pp_rrt.cu

struct pp_handle {
  pp_params P;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;      // compute
  cudaStream_t copy_stream = nullptr; // H2D of the host-pointer entry points
  cudaStream_t out_stream = nullptr;  // D2H of the host-pointer entry points
  cudaEvent_t* chunk_ev = nullptr;    // 2 events per pipeline chunk (uploaded / computed)
  int chunk_ev_n = 0;
  cudaEvent_t ev[16] = {};
  pp::Lattice lat{};
  pp::GridDev grid{};
  bool has_grid = false;
  // grid bank (pp_set_grid_bank_dev): bank_grid describes grid 0, bank the strides; bank.n == 0: none
  pp::GridDev bank_grid{};
  pp::GridBank bank{};
  int bank_cap = 0;
  uint8_t* bank_rows = nullptr;    // REF_AABB pass-1 scratch of the bank
  int* d_box_tables = nullptr;     // LP:473-476 interval tables (xlo | xhi | ylo | yhi)
  int16_t* d_near_spans = nullptr;   // structuring elements of GridDev::near4: [bin][2 near_e + 1]{first, last} block column
  int near_e = 0;
  int broad_phase = 1;
  bool plan_timing_pending = false;   // ev[8] / ev[9] of an asynchronous pp_plan_batch_dev not read yet
  bool aabb_stream_ready = false;  // dynamic shared-memory limit of collide_aabb_stream_kernel raised on this device
  bool oriented_ready = false;     // dynamic shared-memory size of the ORIENTED batch kernels set on this device
  uint64_t launches = 0;
  float phase_ms[PP_PHASE_COUNT] = {};
  int64_t last_narrow = 0;
  bool narrow_first = true;        // the next ORIENTED launch is the first of a call (see collide_counter_reset)
  // staging for the host-pointer collide entry point
  float* d_poses = nullptr;
  uint8_t* d_flags = nullptr;
  int64_t stage_cap = 0;
  int16_t* d_q16 = nullptr;        // staging of pp_collide_batch_q16
  int64_t q16_cap = 0;
  // compaction scratch for the two-phase ORIENTED path
  int32_t* d_count = nullptr;      // device counter
  int32_t* h_count = nullptr;      // pinned mirror
  void* scratch = nullptr;         // grid-prep scratch
  size_t scratch_bytes = 0;
  uint8_t* d_visited = nullptr;    // A12 debug bitmap of the last single plan (W*H bytes)
  unsigned long long* d_nn_best = nullptr;   // nearest-neighbour merge cells, one per query
  int64_t nn_cap = 0;
  PlanWorkspace* plan_ws = nullptr;
  RrtWorkspace* rrt_ws = nullptr;
};

namespace pp {
inline CollideParams collide_params(const pp_handle* h) {
  CollideParams cp;
  cp.lat = h->lat;
  cp.car_length = h->P.car_length;
  cp.car_width = h->P.car_width;
  cp.search_res = h->P.search_res;
  cp.mode = h->P.collision_mode;
  return cp;
}
// kernels / host helpers implemented across the .cu files
int grid_upload(pp_handle* h, const int8_t* data, bool data_on_device, int layout);
int grid_bank_upload(pp_handle* h, int n, const int8_t* grids_dev, int layout);
void grid_free(pp_handle* h);
int collide_batch_dev(pp_handle* h, int64_t n, const float* poses_dev, uint8_t* flags_dev, bool bits_out = false);
int collide_launch(pp_handle* h, int64_t n, const float* poses_dev, uint8_t* flags_dev, bool bits_out = false);   // bits_out: ORIENTED only
int collide_counter_reset(pp_handle* h);
int collide_counter_fetch(pp_handle* h);
int collide_edges_dev(pp_handle* h, int64_t n, const float* edges_dev, uint8_t* flags_dev);
int pack_flags_dev(pp_handle* h, int64_t n, const uint8_t* flags_dev, uint32_t* bits_dev);
int dequant_poses_dev(pp_handle* h, int64_t n, const int16_t* q_dev, const pp_pose_quant& k, float* poses_dev);
int unpack_flags_dev(pp_handle* h, int64_t n, const uint32_t* bits_dev, uint8_t* flags_dev);
int footprint_points_dev(pp_handle* h, double x, double y, double yaw, double* out_dev, int cap, int* n);
int sample_poses_dev(pp_handle* h, uint64_t seed, uint32_t stream, uint64_t first, int64_t n, float x_lo,
                     float x_hi, float y_lo, float y_hi, float* out);
int nearest_dev(pp_handle* h, int64_t n_nodes, const float* node_xy, int64_t n_query, const float* q,
                int32_t* idx, float* d2);
int l2_gather_bench(pp_handle* h, size_t buffer_bytes, double* out_gbps);
int find_exact_dev(pp_handle* h, int64_t n_nodes, const double* keys, int64_t n_query, const double* q,
                   int32_t* idx);
int plan_batch(pp_handle* h, int n, const pp_query* qs, pp_result* rs, bool banked = false, unsigned char* rec_dev = nullptr,
               int rec_cap_path = 0);
int rrt_plan_batch(pp_handle* h, int n, const pp_query* qs, pp_result* rs);
void plan_free(pp_handle* h);
void rrt_free(pp_handle* h);
Lattice make_lattice(double car_length, double car_width, double res);
bool lattice_supported(const Lattice& l);
}  // namespace pp
