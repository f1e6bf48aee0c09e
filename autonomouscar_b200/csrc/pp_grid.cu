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

namespace pp {

namespace {

// LP:437-441: location = H*W - i*H - j - 1 (exact in double for any grid that fits memory)
__global__ void flip_msg_to_cspace(const int8_t* __restrict__ msg, int8_t* __restrict__ cspace, int W, int H) {
  const int64_t total = static_cast<int64_t>(W) * H;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // k = i*H + j  ->  location = total - k - 1
    cspace[k] = msg[total - k - 1];
  }
}

// one warp packs 32 consecutive cells of a row into one word with a ballot
__global__ void pack_occ_tiles(const int8_t* __restrict__ cspace, uint32_t* __restrict__ tiles, int W, int H,
                               int TI, int TJ) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_words = static_cast<int64_t>(TI) * TJ * kTile;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp; w < n_words; w += n_warps) {
    const int r = static_cast<int>(w & 31);
    const int64_t t = w >> 5;
    const int tj = static_cast<int>(t % TJ), ti = static_cast<int>(t / TJ);
    const int i = ti * kTile - kHaloCells + r;
    const int j = tj * kTile - kHaloCells + lane;
    bool occ = false;
    if (i >= 0 && i < W && j >= 0 && j < H) occ = cspace[static_cast<size_t>(i) * H + j] == 100;
    const uint32_t word = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) tiles[w] = word;
  }
}

__global__ void reduce_tile_any(const uint32_t* __restrict__ tiles, uint8_t* __restrict__ tile_any, int64_t n_tiles) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    const uint32_t w = tiles[t * kTile + lane];
    const bool any = __any_sync(0xffffffffu, w != 0);
    if (lane == 0) tile_any[t] = any ? 1 : 0;
  }
}

// 8x8-cell coarse occupancy: thread per (coarse row, tile column) ORs 8 tile words -> 4 bits
__global__ void build_coarse(const uint32_t* __restrict__ tiles, unsigned long long* __restrict__ coarse, int TI,
                             int TJ, int CW8) {
  const int64_t total = static_cast<int64_t>(TI) * 4 * TJ;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tj = static_cast<int>(k % TJ);
    const int ci8 = static_cast<int>(k / TJ);
    const int ti = ci8 >> 2, r0 = (ci8 & 3) * 8;
    const uint32_t* t = tiles + (static_cast<size_t>(ti) * TJ + tj) * kTile + r0;
    uint32_t w = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) w |= t[r];
    unsigned long long bits = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if ((w >> (8 * b)) & 0xffu) bits |= 1ull << b;
    if (bits) {
      const int cj8 = tj * 4;
      atomicOr(&coarse[static_cast<size_t>(ci8) * CW8 + (cj8 >> 6)], bits << (cj8 & 63));
    }
  }
}

// 4x4-cell coarse occupancy: thread per (coarse row, tile column) ORs 4 tile words -> 8 bits
__global__ void build_coarse4(const uint32_t* __restrict__ tiles, unsigned long long* __restrict__ coarse4, int TI,
                              int TJ, int CW4) {
  const int64_t total = static_cast<int64_t>(TI) * 8 * TJ;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tj = static_cast<int>(k % TJ);
    const int ci4 = static_cast<int>(k / TJ);
    const int ti = ci4 >> 3, r0 = (ci4 & 7) * 4;
    const uint32_t* t = tiles + (static_cast<size_t>(ti) * TJ + tj) * kTile + r0;
    const uint32_t w = t[0] | t[1] | t[2] | t[3];
    unsigned long long bits = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if ((w >> (4 * b)) & 0xfu) bits |= 1ull << b;
    if (bits) {
      const int cj4 = tj * 8;
      atomicOr(&coarse4[static_cast<size_t>(ci4) * CW4 + (cj4 >> 6)], bits << (cj4 & 63));
    }
  }
}

// near8 = coarse dilated by r8 blocks (square structuring element), one thread per 64-bit word
__global__ void dilate_coarse(const unsigned long long* __restrict__ coarse, unsigned long long* __restrict__ near8,
                              int CI8, int CW8, int r8) {
  const int64_t total = static_cast<int64_t>(CI8) * CW8;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(k % CW8), r = static_cast<int>(k / CW8);
    unsigned long long acc = 0;
    for (int rr = max(r - r8, 0); rr <= min(r + r8, CI8 - 1); ++rr) {
      const unsigned long long* row = coarse + static_cast<size_t>(rr) * CW8;
      const unsigned long long c = row[w], l = w > 0 ? row[w - 1] : 0ull, h = w + 1 < CW8 ? row[w + 1] : 0ull;
      unsigned long long d = c;
      for (int sft = 1; sft <= r8; ++sft) {        // r8 < 64
        d |= (c << sft) | (l >> (64 - sft));       // bits moving up from lower columns
        d |= (c >> sft) | (h << (64 - sft));       // bits moving down from higher columns
      }
      acc |= d;
    }
    near8[k] = acc;
  }
}

// REF_AABB dilation, pass 1: rows[x][my] = OR over y in [ylo[my], yhi[my]) of (cspace[x][y]==100)
__global__ void aabb_pass_y(const int8_t* __restrict__ cspace, uint8_t* __restrict__ rows, int W, int H, int ny,
                            const int* __restrict__ ylo, const int* __restrict__ yhi) {
  const int64_t total = static_cast<int64_t>(W) * ny;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int my = static_cast<int>(k % ny);
    const int x = static_cast<int>(k / ny);
    int lo = max(ylo[my], 0), hi = min(yhi[my], H);
    uint8_t any = 0;
    for (int y = lo; y < hi; ++y) any |= (cspace[static_cast<size_t>(x) * H + y] == 100);
    rows[k] = any;
  }
}
// pass 2: map[mx][my] = OR over x in [xlo[mx], xhi[mx]) of rows[x][my]
__global__ void aabb_pass_x(const uint8_t* __restrict__ rows, uint8_t* __restrict__ map, int W, int nx, int ny,
                            const int* __restrict__ xlo, const int* __restrict__ xhi) {
  const int64_t total = static_cast<int64_t>(nx) * ny;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int my = static_cast<int>(k % ny);
    const int mx = static_cast<int>(k / ny);
    int lo = max(xlo[mx], 0), hi = min(xhi[mx], W);
    uint8_t any = 0;
    for (int x = lo; x < hi; ++x) any |= rows[static_cast<size_t>(x) * ny + my];
    map[k] = any;
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

int blocks_for(int64_t work, int threads, int sm_count) {
  const int64_t want = (work + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count) * 8;
  return static_cast<int>(std::max<int64_t>(1, std::min(want, cap)));
}

}  // namespace

void grid_free(pp_handle* h) {
  GridDev& g = h->grid;
  cudaFree(g.cspace);
  cudaFree(g.occ_tiles);
  cudaFree(g.tile_any);
  cudaFree(g.coarse);
  cudaFree(g.coarse4);
  cudaFree(g.near8);
  cudaFree(g.aabb_map);
  g = GridDev{};
  h->has_grid = false;
}

int grid_upload(pp_handle* h, const int8_t* data, bool data_on_device, int layout) {
  PP_CUDA(cudaSetDevice(h->device));
  GridDev& g = h->grid;
  const int W = h->P.costmap_width, H = h->P.costmap_height;
  const size_t cells = static_cast<size_t>(W) * H;
  if (!g.cspace) {  // first grid: allocate every representation once
    g.W = W; g.H = H; g.res = h->P.costmap_res;
    g.TI = (W + 2 * kHaloCells + kTile - 1) / kTile;
    g.TJ = (H + 2 * kHaloCells + kTile - 1) / kTile;
    g.aabb_pad_x = static_cast<int>(std::ceil(h->P.car_length / 2 / g.res)) + 2;
    g.aabb_pad_y = static_cast<int>(std::ceil(h->P.car_width / 2 / g.res)) + 2;
    g.aabb_nx = W + 2 * g.aabb_pad_x;
    g.aabb_ny = H + 2 * g.aabb_pad_y;
    PP_CUDA(cudaMalloc(&g.cspace, cells));
    // + 2 tiles of padding: the narrow phase reads every tile row three tiles wide
    PP_CUDA(cudaMalloc(&g.occ_tiles, (static_cast<size_t>(g.TI) * g.TJ + 2) * kTile * sizeof(uint32_t)));
    PP_CUDA(cudaMemset(g.occ_tiles + static_cast<size_t>(g.TI) * g.TJ * kTile, 0, 2 * kTile * sizeof(uint32_t)));
    PP_CUDA(cudaMalloc(&g.tile_any, static_cast<size_t>(g.TI) * g.TJ));
    g.CI8 = g.TI * 4;
    g.CW8 = (g.TJ * 4 + 63) / 64 + 1;  // +1: the bbox test may read one word past the last
    PP_CUDA(cudaMalloc(&g.coarse, sizeof(unsigned long long) * static_cast<size_t>(g.CI8) * g.CW8));
    PP_CUDA(cudaMalloc(&g.near8, sizeof(unsigned long long) * static_cast<size_t>(g.CI8) * g.CW8));
    g.CI4 = g.TI * 8;
    g.CW4 = (g.TJ * 8 + 63) / 64 + 1;
    PP_CUDA(cudaMalloc(&g.coarse4, sizeof(unsigned long long) * static_cast<size_t>(g.CI4) * g.CW4));
    {
      // half diagonal of the footprint in cells + 1 (floor) + 2 (float slack of the pre-test)
      const double du = h->lat.su * h->lat.hu, dv = h->lat.sv * h->lat.hv;
      g.reach_cells = static_cast<int>(std::sqrt(du * du + dv * dv)) + 4;
    }
    PP_CUDA(cudaMalloc(&g.aabb_map, static_cast<size_t>(g.aabb_nx) * g.aabb_ny));
    // scratch: message copy (cells) + pass-1 rows (W * ny) + 4 interval tables
    h->scratch_bytes = cells + static_cast<size_t>(W) * g.aabb_ny + sizeof(int) * 2 * (g.aabb_nx + g.aabb_ny) + 256;
    PP_CUDA(cudaMalloc(&h->scratch, h->scratch_bytes));
    // interval tables depend only on the parameters: upload once
    std::vector<int> xlo, xhi, ylo, yhi;
    box_interval_table(W, g.aabb_pad_x, h->P.car_length, g.res, &xlo, &xhi);
    box_interval_table(H, g.aabb_pad_y, h->P.car_width, g.res, &ylo, &yhi);
    char* base = static_cast<char*>(h->scratch) + ((cells + static_cast<size_t>(W) * g.aabb_ny + 255) & ~size_t(255));
    int* t = reinterpret_cast<int*>(base);
    PP_CUDA(cudaMemcpy(t, xlo.data(), sizeof(int) * g.aabb_nx, cudaMemcpyHostToDevice));
    PP_CUDA(cudaMemcpy(t + g.aabb_nx, xhi.data(), sizeof(int) * g.aabb_nx, cudaMemcpyHostToDevice));
    PP_CUDA(cudaMemcpy(t + 2 * g.aabb_nx, ylo.data(), sizeof(int) * g.aabb_ny, cudaMemcpyHostToDevice));
    PP_CUDA(cudaMemcpy(t + 2 * g.aabb_nx + g.aabb_ny, yhi.data(), sizeof(int) * g.aabb_ny, cudaMemcpyHostToDevice));
  }
  int8_t* msg = static_cast<int8_t*>(h->scratch);
  uint8_t* rows = reinterpret_cast<uint8_t*>(h->scratch) + cells;
  char* tbase = static_cast<char*>(h->scratch) + ((cells + static_cast<size_t>(W) * g.aabb_ny + 255) & ~size_t(255));
  const int* xlo = reinterpret_cast<const int*>(tbase);
  const int* xhi = xlo + g.aabb_nx;
  const int* ylo = xhi + g.aabb_nx;
  const int* yhi = ylo + g.aabb_ny;

  cudaStream_t s = h->stream;
  const int T = 256;
  if (layout == PP_GRID_OCCUPANCY_MSG) {
    const int8_t* src = data;
    if (!data_on_device) {
      PP_CUDA(cudaMemcpyAsync(msg, data, cells, cudaMemcpyHostToDevice, s));
      src = msg;
    }
    flip_msg_to_cspace<<<blocks_for(cells, T, h->sm_count), T, 0, s>>>(src, g.cspace, W, H);
    h->launches++;
  } else {
    PP_CUDA(cudaMemcpyAsync(g.cspace, data, cells, data_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  }
  const int64_t n_tiles = static_cast<int64_t>(g.TI) * g.TJ;
  pack_occ_tiles<<<blocks_for(n_tiles * kTile * 32, T, h->sm_count), T, 0, s>>>(g.cspace, g.occ_tiles, W, H, g.TI, g.TJ);
  reduce_tile_any<<<blocks_for(n_tiles * 32, T, h->sm_count), T, 0, s>>>(g.occ_tiles, g.tile_any, n_tiles);
  PP_CUDA(cudaMemsetAsync(g.coarse, 0, sizeof(unsigned long long) * static_cast<size_t>(g.CI8) * g.CW8, s));
  build_coarse<<<blocks_for(static_cast<int64_t>(g.CI8) * g.TJ, T, h->sm_count), T, 0, s>>>(g.occ_tiles, g.coarse, g.TI, g.TJ, g.CW8);
  dilate_coarse<<<blocks_for(static_cast<int64_t>(g.CI8) * g.CW8, T, h->sm_count), T, 0, s>>>(g.coarse, g.near8, g.CI8, g.CW8, (g.reach_cells + 7) / 8);
  PP_CUDA(cudaMemsetAsync(g.coarse4, 0, sizeof(unsigned long long) * static_cast<size_t>(g.CI4) * g.CW4, s));
  build_coarse4<<<blocks_for(static_cast<int64_t>(g.CI4) * g.TJ, T, h->sm_count), T, 0, s>>>(g.occ_tiles, g.coarse4, g.TI, g.TJ, g.CW4);
  h->launches += 3;
  aabb_pass_y<<<blocks_for(static_cast<int64_t>(W) * g.aabb_ny, T, h->sm_count), T, 0, s>>>(g.cspace, rows, W, H, g.aabb_ny, ylo, yhi);
  aabb_pass_x<<<blocks_for(static_cast<int64_t>(g.aabb_nx) * g.aabb_ny, T, h->sm_count), T, 0, s>>>(rows, g.aabb_map, W, g.aabb_nx, g.aabb_ny, xlo, xhi);
  h->launches += 4;
  PP_CUDA(cudaGetLastError());
  PP_CUDA(cudaStreamSynchronize(s));
  h->has_grid = true;
  return PP_OK;
}

}  // namespace pp
