// pp_grid.cu -- grid ingest: replaces costmapCallback -> mapCostmapToCSpace ->
// calcGridLocation (local_planner.cpp:658-661, 425-441) and builds the device-side
// representations the collision kernels gather from.
//
//   cspace     W*H int8, the planner's m_c_space order              (HBM: 1 B / cell)
//   occ_tiles  (cell == 100) bit-packed in 32x32-cell tiles, zero halo of 128 cells;
//              one tile = 32 words = exactly one 128 B line, so a footprint window
//              (<= 52 x 52 cells) is at most 3 x 3 lines                (HBM: 1 bit / cell)
//   tile_any   1 B per tile, broad-phase "tile has an obstacle"
//   aabb_map   the reference's axis-aligned cell box (LP:467-496) OR-reduced per base
//              cell, so REF_AABB is one lookup per pose
#include <algorithm>
#include <cmath>
#include <vector>

#include "pp_common.cuh"
#include "pp_near_spans.h"

namespace pp {

namespace {

// Every build kernel handles grid blockIdx.y of a bank (grids of one shape stored back to back, pp_common.cuh);
// the planner's own single grid is a bank of one with zero strides.
struct BuildArgs {
  GridDev g;                 // grid 0 of the bank
  GridBank b;
  const int8_t* src;         // OccupancyGrid.data of grid 0 (message order), src_stride cells apart
  size_t src_stride;
  uint8_t* rows;             // REF_AABB pass-1 scratch of grid 0, rows_stride bytes apart
  size_t rows_stride;
  const int *xlo, *xhi, *ylo, *yhi;
  const int16_t* near_spans;  // [bin][2 near_e + 1]{first, last}: structuring elements of near4
  int near_e;
};
__device__ __forceinline__ GridDev build_grid(const BuildArgs& a) { return grid_at(a.g, a.b, blockIdx.y); }

// LP:437-441: location = H*W - i*H - j - 1 (exact in double for any grid that fits memory)
__global__ void flip_msg_to_cspace(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int8_t* __restrict__ msg = a.src + blockIdx.y * a.src_stride;
  const int64_t total = static_cast<int64_t>(g.W) * g.H;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // k = i*H + j  ->  location = total - k - 1
    g.cspace[k] = msg[total - k - 1];
  }
}

// one warp packs 32 consecutive cells of a row into one word with a ballot
__global__ void pack_occ_tiles(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int W = g.W, H = g.H, TJ = g.TJ;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_words = static_cast<int64_t>(g.TI) * TJ * kTile;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp; w < n_words; w += n_warps) {
    const int r = static_cast<int>(w & 31);
    const int64_t t = w >> 5;
    const int tj = static_cast<int>(t % TJ), ti = static_cast<int>(t / TJ);
    const int i = ti * kTile - kHaloCells + r;
    const int j = tj * kTile - kHaloCells + lane;
    bool occ = false;
    if (i >= 0 && i < W && j >= 0 && j < H) occ = g.cspace[static_cast<size_t>(i) * H + j] == 100;
    const uint32_t word = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) g.occ_tiles[w] = word;
  }
}

__global__ void reduce_tile_any(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int64_t n_tiles = static_cast<int64_t>(g.TI) * g.TJ;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    const uint32_t w = g.occ_tiles[t * kTile + lane];
    const bool any = __any_sync(0xffffffffu, w != 0);
    if (lane == 0) g.tile_any[t] = any ? 1 : 0;
  }
}

// 8x8-cell coarse occupancy: thread per (coarse row, tile column) ORs 8 tile words -> 4 bits
__global__ void build_coarse(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int TJ = g.TJ;
  const int64_t total = static_cast<int64_t>(g.TI) * 4 * TJ;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tj = static_cast<int>(k % TJ);
    const int ci8 = static_cast<int>(k / TJ);
    const int ti = ci8 >> 2, r0 = (ci8 & 3) * 8;
    const uint32_t* t = g.occ_tiles + (static_cast<size_t>(ti) * TJ + tj) * kTile + r0;
    uint32_t w = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) w |= t[r];
    unsigned long long bits = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if ((w >> (8 * b)) & 0xffu) bits |= 1ull << b;
    if (bits) {
      const int cj8 = tj * 4;
      atomicOr(&g.coarse[static_cast<size_t>(ci8) * g.CW8 + (cj8 >> 6)], bits << (cj8 & 63));
    }
  }
}

// 4x4-cell coarse occupancy ("any" and "full"): thread per (coarse row, tile column) combines 4 tile words -> 8 bits
__global__ void build_coarse4(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int TJ = g.TJ;
  const int64_t total = static_cast<int64_t>(g.TI) * 8 * TJ;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tj = static_cast<int>(k % TJ);
    const int ci4 = static_cast<int>(k / TJ);
    const int ti = ci4 >> 3, r0 = (ci4 & 7) * 4;
    const uint32_t* t = g.occ_tiles + (static_cast<size_t>(ti) * TJ + tj) * kTile + r0;
    const uint32_t w = t[0] | t[1] | t[2] | t[3], wa = t[0] & t[1] & t[2] & t[3];
    unsigned long long bits = 0, full = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      if ((w >> (4 * b)) & 0xfu) bits |= 1ull << b;
      if (((wa >> (4 * b)) & 0xfu) == 0xfu) full |= 1ull << b;
    }
    const int cj4 = tj * 8;
    // word pairs: {any, full}; "full" = all 16 cells occupied (coarse4_oriented_test's collision proof)
    const size_t at = 2 * (static_cast<size_t>(ci4) * g.CW4 + (cj4 >> 6));
    if (bits) atomicOr(&g.coarse4[at], bits << (cj4 & 63));
    if (full) atomicOr(&g.coarse4[at + 1], full << (cj4 & 63));
  }
}

// near4[bin] = the 4x4-block "any" map dilated by the bin's structuring element: output block (r, c) is set iff an
// input block (r + di, c + dj) is, for some row di in [-near_e, near_e] and dj in that row's span.  One thread per
// 64-bit output word.
__global__ void dilate_near4(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int CI4 = g.CI4, CW4 = g.CW4, E = a.near_e;
  const int64_t per_bin = static_cast<int64_t>(CI4) * CW4, total = per_bin * kNearBins;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int bin = static_cast<int>(k / per_bin);
    const int64_t kk = k - bin * per_bin;
    const int w = static_cast<int>(kk % CW4), r = static_cast<int>(kk / CW4);
    const int16_t* spans = a.near_spans + static_cast<size_t>(bin) * (2 * E + 1) * 2;
    unsigned long long acc = 0;
    for (int di = -E; di <= E; ++di) {
      const int rr = r + di, lo = spans[2 * (di + E)], hi = spans[2 * (di + E) + 1];   // |lo|, |hi| < 64
      if (rr < 0 || rr >= CI4 || lo > hi) continue;
      const unsigned long long* row = g.coarse4 + 2 * static_cast<size_t>(rr) * CW4;   // ("any" words of the pairs)
      const unsigned long long c = row[2 * w], l = w > 0 ? row[2 * (w - 1)] : 0ull, h = w + 1 < CW4 ? row[2 * (w + 1)] : 0ull;
      for (int dj = lo; dj <= hi; ++dj) {          // output bit p takes input bit p + dj
        if (dj == 0) acc |= c;
        else if (dj > 0) acc |= (c >> dj) | (h << (64 - dj));
        else acc |= (c << -dj) | (l >> (64 + dj));
      }
    }
    g.near4[k] = acc;
  }
}

// REF_AABB dilation, pass 1: rows[x][my] = OR over y in [ylo[my], yhi[my]) of (cspace[x][y]==100)
__global__ void aabb_pass_y(BuildArgs a) {
  const GridDev g = build_grid(a);
  uint8_t* __restrict__ rows = a.rows + blockIdx.y * a.rows_stride;
  const int H = g.H, ny = g.aabb_ny;
  const int64_t total = static_cast<int64_t>(g.W) * ny;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int my = static_cast<int>(k % ny);
    const int x = static_cast<int>(k / ny);
    int lo = max(a.ylo[my], 0), hi = min(a.yhi[my], H);
    uint8_t any = 0;
    for (int y = lo; y < hi; ++y) any |= (g.cspace[static_cast<size_t>(x) * H + y] == 100);
    rows[k] = any;
  }
}
// pass 2: map[mx][my] = OR over x in [xlo[mx], xhi[mx]) of rows[x][my]
__global__ void aabb_pass_x(BuildArgs a) {
  const GridDev g = build_grid(a);
  const uint8_t* __restrict__ rows = a.rows + blockIdx.y * a.rows_stride;
  const int W = g.W, ny = g.aabb_ny;
  const int64_t total = static_cast<int64_t>(g.aabb_nx) * ny;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int my = static_cast<int>(k % ny);
    const int mx = static_cast<int>(k / ny);
    int lo = max(a.xlo[mx], 0), hi = min(a.xhi[mx], W);
    uint8_t any = 0;
    for (int x = lo; x < hi; ++x) any |= rows[static_cast<size_t>(x) * ny + my];
    g.aabb_map[k] = any;
  }
}

// 2-bit summary of aabb_map per block of (1 << s)^2 map cells (pp_common.cuh): one thread per block
__global__ void aabb_coarse_kernel(BuildArgs a) {
  const GridDev g = build_grid(a);
  const int s = g.aabb_cshift, B = 1 << s, nby = (g.aabb_ny + B - 1) >> s;
  const int64_t total = static_cast<int64_t>(g.aabb_cnbx) * nby;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cy = static_cast<int>(k % nby), cx = static_cast<int>(k / nby);
    const int x0 = cx << s, y0 = cy << s;
    const int x1 = min(x0 + B, g.aabb_nx), y1 = min(y0 + B, g.aabb_ny);
    unsigned any = 0, all = (x1 - x0 == B && y1 - y0 == B) ? 1u : 0u;
    for (int x = x0; x < x1; ++x)
      for (int y = y0; y < y1; ++y) {
        const unsigned v = g.aabb_map[static_cast<size_t>(x) * g.aabb_ny + y] != 0;
        any |= v;
        all &= v;
      }
    const unsigned bits = any | (all << 1);
    if (bits) atomicOr(&g.aabb_coarse[static_cast<size_t>(cx) * g.aabb_cwpr + (cy >> 4)], bits << ((cy & 15) * 2));
  }
}

// LP:473-476 evaluated on the host in double exactly as the reference would:
//   lo = int(base - extent/2/res), hi = int(base + extent/2/res)  (truncation toward zero)
void box_interval_table(int n_cells, int pad, double extent, double res, std::vector<int>* lo,
                        std::vector<int>* hi) {
  const int n = n_cells + 2 * pad;
  lo->resize(n);
  hi->resize(n);
  for (int m = 0; m < n; ++m) {
    const int base = m - pad;
    (*lo)[m] = static_cast<int>(base - extent / 2 / res);
    (*hi)[m] = static_cast<int>(base + extent / 2 / res);
  }
}

int blocks_for(int64_t work, int threads, int sm_count, int n_grids = 1) {
  const int64_t want = (work + threads - 1) / threads;
  const int64_t cap = std::max<int64_t>(4, static_cast<int64_t>(sm_count) * 8 / n_grids);
  return static_cast<int>(std::max<int64_t>(1, std::min(want, cap)));
}

// shapes of every representation for the handle's parameters; pointers left null
GridDev grid_shape(const pp_handle* h) {
  GridDev g{};
  const int W = h->P.costmap_width, H = h->P.costmap_height;
  g.W = W; g.H = H; g.res = h->P.costmap_res;
  g.TI = (W + 2 * kHaloCells + kTile - 1) / kTile;
  g.TJ = (H + 2 * kHaloCells + kTile - 1) / kTile;
  g.aabb_pad_x = static_cast<int>(std::ceil(h->P.car_length / 2 / g.res)) + 2;
  g.aabb_pad_y = static_cast<int>(std::ceil(h->P.car_width / 2 / g.res)) + 2;
  g.aabb_nx = W + 2 * g.aabb_pad_x;
  g.aabb_ny = H + 2 * g.aabb_pad_y;
  g.CI8 = g.TI * 4;
  g.CW8 = (g.TJ * 4 + 63) / 64 + 1;  // +1: the bbox test may read one word past the last
  g.CI4 = g.TI * 8;
  g.CW4 = (g.TJ * 8 + 63) / 64 + 1;
  // block size of the 2-bit summary: the smallest (>= 16 cells) whose map fits 24 KB, so that two CTAs of the
  // streaming kernel share an SM (4096^2: 16-cell blocks, 17.7 KB, ~10 % of the poses need the cell look-up)
  for (g.aabb_cshift = 4;; ++g.aabb_cshift) {
    const int B = 1 << g.aabb_cshift;
    g.aabb_cnbx = (g.aabb_nx + B - 1) / B;
    g.aabb_cwpr = ((g.aabb_ny + B - 1) / B + 15) / 16;
    if (static_cast<size_t>(g.aabb_cnbx) * g.aabb_cwpr * sizeof(uint32_t) <= 24 * 1024) break;
  }
  // half diagonal of the footprint in cells + 1 (floor) + 2 (float slack of the pre-test)
  const double du = h->lat.su * h->lat.hu, dv = h->lat.sv * h->lat.hv;
  g.reach_cells = static_cast<int>(std::sqrt(du * du + dv * dv)) + 4;
  return g;
}

// element strides between consecutive grids of a bank (16-byte aligned sub-arrays)
GridBank bank_strides(const GridDev& g, int n) {
  auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
  GridBank b;
  b.n = n;
  b.cspace = up(static_cast<size_t>(g.W) * g.H, 16);
  // + 2 tiles of padding: the narrow phase reads every tile row three tiles wide
  b.occ_tiles = (static_cast<size_t>(g.TI) * g.TJ + 2) * kTile;
  b.tile_any = up(static_cast<size_t>(g.TI) * g.TJ, 16);
  b.coarse = static_cast<size_t>(g.CI8) * g.CW8;
  b.near4 = static_cast<size_t>(kNearBins) * g.CI4 * g.CW4;
  b.coarse4 = 2 * static_cast<size_t>(g.CI4) * g.CW4;               // {any, full} word pairs
  b.aabb_map = up(static_cast<size_t>(g.aabb_nx) * g.aabb_ny, 16);
  b.aabb_coarse = up(static_cast<size_t>(g.aabb_cnbx) * g.aabb_cwpr, 4);
  return b;
}

void grid_free_arrays(GridDev* g) {
  cudaFree(g->cspace);
  cudaFree(g->occ_tiles);
  cudaFree(g->tile_any);
  cudaFree(g->coarse);
  cudaFree(g->coarse4);
  cudaFree(g->near4);
  cudaFree(g->aabb_map);
  cudaFree(g->aabb_coarse);
  *g = GridDev{};
}

int grid_alloc(GridDev* g, const GridBank& b) {
  const size_t n = static_cast<size_t>(b.n);
  PP_CUDA(cudaMalloc(&g->cspace, b.cspace * n));
  PP_CUDA(cudaMalloc(&g->occ_tiles, b.occ_tiles * n * sizeof(uint32_t)));
  PP_CUDA(cudaMemset(g->occ_tiles, 0, b.occ_tiles * n * sizeof(uint32_t)));      // (the padding tiles stay zero)
  PP_CUDA(cudaMalloc(&g->tile_any, b.tile_any * n));
  PP_CUDA(cudaMalloc(&g->coarse, sizeof(unsigned long long) * b.coarse * n));
  PP_CUDA(cudaMalloc(&g->near4, sizeof(unsigned long long) * b.near4 * n));
  PP_CUDA(cudaMalloc(&g->coarse4, sizeof(unsigned long long) * b.coarse4 * n));
  PP_CUDA(cudaMalloc(&g->aabb_map, b.aabb_map * n));
  PP_CUDA(cudaMalloc(&g->aabb_coarse, sizeof(uint32_t) * b.aabb_coarse * n));
  return PP_OK;
}

int ensure_near_spans(pp_handle* h) {
  if (h->d_near_spans) return PP_OK;
  std::vector<int16_t> t;
  near_span_table(h->lat.su * h->lat.hu, h->lat.sv * h->lat.hv, kNearBins, &h->near_e, &t);
  if (h->near_e >= 60) return PP_ERR_UNSUPPORTED;   // (spans must stay inside the neighbouring 64-bit words)
  PP_CUDA(cudaMalloc(&h->d_near_spans, sizeof(int16_t) * t.size()));
  PP_CUDA(cudaMemcpy(h->d_near_spans, t.data(), sizeof(int16_t) * t.size(), cudaMemcpyHostToDevice));
  return PP_OK;
}

// the LP:473-476 interval tables depend only on the parameters: built once per handle
int ensure_box_tables(pp_handle* h, const GridDev& g) {
  if (h->d_box_tables) return PP_OK;
  std::vector<int> xlo, xhi, ylo, yhi;
  box_interval_table(g.W, g.aabb_pad_x, h->P.car_length, g.res, &xlo, &xhi);
  box_interval_table(g.H, g.aabb_pad_y, h->P.car_width, g.res, &ylo, &yhi);
  PP_CUDA(cudaMalloc(&h->d_box_tables, sizeof(int) * 2 * (g.aabb_nx + g.aabb_ny)));
  int* t = h->d_box_tables;
  PP_CUDA(cudaMemcpy(t, xlo.data(), sizeof(int) * g.aabb_nx, cudaMemcpyHostToDevice));
  PP_CUDA(cudaMemcpy(t + g.aabb_nx, xhi.data(), sizeof(int) * g.aabb_nx, cudaMemcpyHostToDevice));
  PP_CUDA(cudaMemcpy(t + 2 * g.aabb_nx, ylo.data(), sizeof(int) * g.aabb_ny, cudaMemcpyHostToDevice));
  PP_CUDA(cudaMemcpy(t + 2 * g.aabb_nx + g.aabb_ny, yhi.data(), sizeof(int) * g.aabb_ny, cudaMemcpyHostToDevice));
  return PP_OK;
}

// all derived representations of the n grids of a bank from their cspace (or message) bytes, on h->stream
int grid_build(pp_handle* h, const GridDev& g, const GridBank& b, const int8_t* msg_src, size_t src_stride, uint8_t* rows,
               size_t rows_stride) {
  cudaStream_t s = h->stream;
  const int T = 256, n = b.n;
  const unsigned ny = static_cast<unsigned>(n);
  BuildArgs a;
  a.g = g; a.b = b; a.src = msg_src; a.src_stride = src_stride; a.rows = rows; a.rows_stride = rows_stride;
  a.xlo = h->d_box_tables; a.xhi = a.xlo + g.aabb_nx; a.ylo = a.xhi + g.aabb_nx; a.yhi = a.ylo + g.aabb_ny;
  const int nrc = ensure_near_spans(h);
  if (nrc) return nrc;
  a.near_spans = h->d_near_spans; a.near_e = h->near_e;
  auto blocks = [&](int64_t work) { return dim3(static_cast<unsigned>(blocks_for(work, T, h->sm_count, n)), ny); };
  const int64_t cells = static_cast<int64_t>(g.W) * g.H;
  if (msg_src) {
    flip_msg_to_cspace<<<blocks(cells), T, 0, s>>>(a);
    h->launches++;
  }
  const int64_t n_tiles = static_cast<int64_t>(g.TI) * g.TJ;
  pack_occ_tiles<<<blocks(n_tiles * kTile * 32), T, 0, s>>>(a);
  reduce_tile_any<<<blocks(n_tiles * 32), T, 0, s>>>(a);
  // (the planner's own grid is a bank of one with ZERO strides: sizes come from the shape then)
  const size_t coarse_n = n > 1 ? b.coarse * n : static_cast<size_t>(g.CI8) * g.CW8;
  const size_t coarse4_n = n > 1 ? b.coarse4 * n : 2 * static_cast<size_t>(g.CI4) * g.CW4;
  const size_t acoarse_n = n > 1 ? b.aabb_coarse * n : static_cast<size_t>(g.aabb_cnbx) * g.aabb_cwpr;
  PP_CUDA(cudaMemsetAsync(g.coarse, 0, sizeof(unsigned long long) * coarse_n, s));
  build_coarse<<<blocks(static_cast<int64_t>(g.CI8) * g.TJ), T, 0, s>>>(a);
  PP_CUDA(cudaMemsetAsync(g.coarse4, 0, sizeof(unsigned long long) * coarse4_n, s));
  build_coarse4<<<blocks(static_cast<int64_t>(g.CI4) * g.TJ), T, 0, s>>>(a);
  dilate_near4<<<blocks(static_cast<int64_t>(kNearBins) * g.CI4 * g.CW4), T, 0, s>>>(a);
  h->launches += 3;
  aabb_pass_y<<<blocks(static_cast<int64_t>(g.W) * g.aabb_ny), T, 0, s>>>(a);
  aabb_pass_x<<<blocks(static_cast<int64_t>(g.aabb_nx) * g.aabb_ny), T, 0, s>>>(a);
  PP_CUDA(cudaMemsetAsync(g.aabb_coarse, 0, sizeof(uint32_t) * acoarse_n, s));
  aabb_coarse_kernel<<<blocks(static_cast<int64_t>(g.aabb_cnbx) * ((g.aabb_ny >> g.aabb_cshift) + 1)), T, 0, s>>>(a);
  h->launches += 5;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}

}  // namespace

void grid_free(pp_handle* h) {
  grid_free_arrays(&h->grid);
  grid_free_arrays(&h->bank_grid);
  cudaFree(h->bank_rows);
  cudaFree(h->d_box_tables);
  cudaFree(h->d_near_spans);
  h->bank_rows = nullptr; h->d_box_tables = nullptr; h->d_near_spans = nullptr;
  h->bank = GridBank{};
  h->has_grid = false;
}

int grid_upload(pp_handle* h, const int8_t* data, bool data_on_device, int layout) {
  PP_CUDA(cudaSetDevice(h->device));
  GridDev& g = h->grid;
  const int W = h->P.costmap_width, H = h->P.costmap_height;
  const size_t cells = static_cast<size_t>(W) * H;
  if (!g.cspace) {  // first grid: allocate every representation once
    g = grid_shape(h);
    const int rc = grid_alloc(&g, bank_strides(g, 1));
    if (rc) return rc;
    // scratch: message copy (cells) + pass-1 rows (W * ny)
    h->scratch_bytes = cells + static_cast<size_t>(W) * g.aabb_ny + 256;
    PP_CUDA(cudaMalloc(&h->scratch, h->scratch_bytes));
  }
  int rc = ensure_box_tables(h, g);
  if (rc) return rc;
  int8_t* msg = static_cast<int8_t*>(h->scratch);
  uint8_t* rows = reinterpret_cast<uint8_t*>(h->scratch) + cells;
  cudaStream_t s = h->stream;
  const int8_t* src = nullptr;
  if (layout == PP_GRID_OCCUPANCY_MSG) {
    src = data;
    if (!data_on_device) {
      PP_CUDA(cudaMemcpyAsync(msg, data, cells, cudaMemcpyHostToDevice, s));
      src = msg;
    }
  } else {
    PP_CUDA(cudaMemcpyAsync(g.cspace, data, cells, data_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  }
  GridBank one;
  one.n = 1;
  rc = grid_build(h, g, one, src, 0, rows, 0);
  if (rc) return rc;
  PP_CUDA(cudaStreamSynchronize(s));
  h->has_grid = true;
  return PP_OK;
}

// n grids of the handle's shape, already on the device, `cells` bytes apart (e.g. pc_batch_grids_dev)
int grid_bank_upload(pp_handle* h, int n, const int8_t* grids_dev, int layout) {
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  const GridDev shape = grid_shape(h);
  const size_t cells = static_cast<size_t>(shape.W) * shape.H;
  if (n > h->bank_cap) {
    grid_free_arrays(&h->bank_grid);
    cudaFree(h->bank_rows);
    h->bank_rows = nullptr; h->bank_cap = 0; h->bank = GridBank{};
    h->bank_grid = shape;
    const GridBank b = bank_strides(shape, n);
    int rc = grid_alloc(&h->bank_grid, b);
    if (rc) return rc;
    PP_CUDA(cudaMalloc(&h->bank_rows, static_cast<size_t>(shape.W) * shape.aabb_ny * n));
    h->bank_cap = n;
  }
  int rc = ensure_box_tables(h, shape);
  if (rc) return rc;
  GridBank b = bank_strides(shape, n);
  const GridDev& g = h->bank_grid;
  cudaStream_t s = h->stream;
  if (layout != PP_GRID_OCCUPANCY_MSG)
    PP_CUDA(cudaMemcpy2DAsync(g.cspace, b.cspace, grids_dev, cells, cells, static_cast<size_t>(n), cudaMemcpyDeviceToDevice, s));
  rc = grid_build(h, g, b, layout == PP_GRID_OCCUPANCY_MSG ? grids_dev : nullptr, cells, h->bank_rows,
                  static_cast<size_t>(shape.W) * shape.aabb_ny);
  if (rc) return rc;
  PP_CUDA(cudaStreamSynchronize(s));
  h->bank = b;
  return PP_OK;
}

}  // namespace pp
