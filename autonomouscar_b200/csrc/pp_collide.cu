// pp_collide.cu -- batched checkCollision (local_planner.cpp:443-496) for independent poses.
//
//   REF_AABB : one thread per pose, LP:471-476 verbatim + one lookup in the dilated map.
//   ORIENTED : broad phase (one thread per pose: fixed-point lattice set-up, bounding box,
//              "any obstacle in the <= 3x3 tiles under the box" test) followed by a
//              warp-cooperative narrow phase: the warp stages the tiles under the footprint
//              (each tile is one coalesced 128 B line) into shared memory and the 32 lanes
//              split the nu x nv sample lattice between them.
//   With broad_phase = 0 every pose that overlaps the grid takes the narrow phase
//   ("direct" mode, the gather-bound measurement).  Results are identical either way.
#include "pp_common.cuh"

namespace pp {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

__global__ void __launch_bounds__(kThreads)
collide_aabb_kernel(GridDev g, const float* __restrict__ poses, uint8_t* __restrict__ flags, int64_t n) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = poses[3 * k], y = poses[3 * k + 1];
    flags[k] = aabb_hit(g, static_cast<double>(x), static_cast<double>(y)) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads)
fill_zero_kernel(uint8_t* __restrict__ flags, int64_t n) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    flags[k] = 0;
}

// What the narrow phase needs to know about one pose, relative to the corner of the first
// tile under its bounding box (so all lattice coordinates are in [0, 96) cells).
struct NarrowJob {
  int32_t oi, oj, ui, uj, vi, vj;
  uint32_t tiles;   // t0i << 16 | t0j : first tile row / column
  uint32_t box;     // ri0 | ri1 << 7 | rj0 << 14 | rj1 << 21 : lattice bbox relative to the tile corner
  uint32_t occ;     // bit (a*3+b): tile (t0i+a, t0j+b) is under the box AND holds an obstacle
};

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

// same for a wider square (the yaw-independent bound): up to 64 coarse cells wide
__device__ __forceinline__ bool coarse_square_any(const GridDev& g, int i0, int i1, int j0, int j1) {
  const int r0 = (i0 + kHaloCells) >> 3, r1 = (i1 + kHaloCells) >> 3;
  const int c0 = (j0 + kHaloCells) >> 3, c1 = (j1 + kHaloCells) >> 3;
  const int w0 = c0 >> 6, sh = c0 & 63, nb = c1 - c0 + 1;
  const unsigned long long m = nb >= 64 ? ~0ull : ((1ull << nb) - 1ull);
  unsigned long long any = 0;
  for (int r = r0; r <= r1; ++r) {
    const unsigned long long* row = g.coarse + static_cast<size_t>(r) * g.CW8 + w0;
    unsigned long long bits = __ldg(row) >> sh;
    if (sh + nb > 64) bits |= __ldg(row + 1) << (64 - sh);
    any |= bits & m;
  }
  return any != 0;
}

// Broad phase for one pose (one thread): false = certainly collision-free.
__device__ __forceinline__ bool make_job(const GridDev& g, const Lattice& l, const PoseFx& p, int broad_phase,
                                         NarrowJob* job) {
  if (!p.valid) return false;
  int i0, i1, j0, j1;
  lattice_bbox(l, p, &i0, &i1, &j0, &j1);
  if (i1 < 0 || i0 >= g.W || j1 < 0 || j0 >= g.H) return false;  // box entirely off the grid
  const int t0i = (i0 + kHaloCells) >> 5, t0j = (j0 + kHaloCells) >> 5;
  const int nti = ((i1 + kHaloCells) >> 5) - t0i + 1, ntj = ((j1 + kHaloCells) >> 5) - t0j + 1;
  uint32_t occ = 0;
  for (int a = 0; a < nti; ++a)
    for (int b = 0; b < ntj; ++b) {
      const uint32_t any = broad_phase ? __ldg(&g.tile_any[static_cast<size_t>(t0i + a) * g.TJ + (t0j + b)]) : 1u;
      occ |= (any ? 1u : 0u) << (a * 3 + b);
    }
  if (occ == 0) return false;                                            // no obstacle in any tile under the box
  if (broad_phase && !coarse_box_any(g, i0, i1, j0, j1)) return false;   // none in the 8x8 blocks under it
  // re-base the lattice origin from the anchor cell to the first tile's corner
  const int ci0 = t0i * kTile - kHaloCells, cj0 = t0j * kTile - kHaloCells;
  job->oi = p.oi + ((p.ai - ci0) << kFracBits);
  job->oj = p.oj + ((p.aj - cj0) << kFracBits);
  job->ui = p.ui; job->uj = p.uj; job->vi = p.vi; job->vj = p.vj;
  job->tiles = (static_cast<uint32_t>(t0i) << 16) | static_cast<uint32_t>(t0j);
  job->box = static_cast<uint32_t>(i0 - ci0) | (static_cast<uint32_t>(i1 - ci0) << 7) |
             (static_cast<uint32_t>(j0 - cj0) << 14) | (static_cast<uint32_t>(j1 - cj0) << 21);
  job->occ = occ;
  return true;
}

// Window of the warp: 96 rows x 96 columns of occupancy bits as sm[wcol * 96 + row]
// (wcol = column / 32).  Each staged tile is one coalesced 128 B line.
constexpr int kWinWords = 3 * 96;

__device__ __forceinline__ void stage_window(const GridDev& g, const NarrowJob& job, uint32_t* sm, int lane) {
  const int t0i = job.tiles >> 16, t0j = job.tiles & 0xffff;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      uint32_t w = 0;
      if ((job.occ >> (a * 3 + b)) & 1u)
        w = __ldg(&g.occ_tiles[(static_cast<size_t>(t0i + a) * g.TJ + (t0j + b)) * kTile + lane]);
      sm[b * 96 + a * 32 + lane] = w;
    }
  }
  __syncwarp();
}

// exact "any occupied cell inside the lattice's bounding box" on the staged window
__device__ __forceinline__ bool window_box_any(const NarrowJob& job, const uint32_t* sm, int lane) {
  const int ri0 = job.box & 127, ri1 = (job.box >> 7) & 127, rj0 = (job.box >> 14) & 127, rj1 = (job.box >> 21) & 127;
  // 96-bit column mask [rj0, rj1] split over the three words
  uint32_t m[3];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    const int lo = max(rj0 - 32 * w, 0), hi = min(rj1 - 32 * w, 31);
    m[w] = (lo <= hi) ? ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo)) : 0u;
  }
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int r = lane + 32 * k;
    if (r >= ri0 && r <= ri1) any |= (sm[r] & m[0]) | (sm[96 + r] & m[1]) | (sm[192 + r] & m[2]);
  }
  return __any_sync(0xffffffffu, any != 0);
}

// Lattice evaluation with a compile-time shape.  The NU columns (fixed a, b = 0..NV-1) are
// cut into S segments of SEG points; lane l owns segment l % S of columns (l / S) * ROUNDS + c
// in round c, so EVERY round samples the whole length of the car at 1/ROUNDS density and a
// colliding pose is almost always caught by the first one (warp-uniform early exit).
// Everything inside is unrolled; each point costs two IMADs, the window address, one LDS,
// a shift and an OR.  The last segment of a column overlaps its neighbour instead of
// overrunning (a sample evaluated twice cannot change an OR).
template <int NU, int NV, int S>
__device__ __forceinline__ bool lattice_any_fixed(const NarrowJob& job, const uint32_t* sm, int lane) {
  constexpr int SEG = (NV + S - 1) / S;
  constexpr int GROUPS = 32 / S;                          // column groups (lanes / S)
  constexpr int ROUNDS = (NU + GROUPS - 1) / GROUPS;
  const int seg = lane % S, grp = lane / S;
  const int b0 = min(seg * SEG, NV - SEG);
  const int a0 = grp * ROUNDS;
  int32_t bi = job.oi + a0 * job.ui + b0 * job.vi;
  int32_t bj = job.oj + a0 * job.uj + b0 * job.vj;
#pragma unroll
  for (int c = 0; c < ROUNDS; ++c) {
    uint32_t hit = 0;
    if ((GROUPS - 1) * ROUNDS + c < NU || a0 + c < NU) {   // first test is compile-time
#pragma unroll
      for (int t = 0; t < SEG; ++t) {
        const uint32_t fi = static_cast<uint32_t>(bi + t * job.vi);
        const uint32_t fj = static_cast<uint32_t>(bj + t * job.vj);
        const uint32_t w = sm[(fj >> 29) * 96 + (fi >> kFracBits)];
        hit |= w >> ((fj >> kFracBits) & 31u);
      }
    }
    if (__any_sync(0xffffffffu, (hit & 1u) != 0)) return true;   // early exit, warp-uniform
    bi += job.ui;
    bj += job.uj;
  }
  return false;
}

// any lattice shape (run-time nu, nv): lane-strided walk in (a, b) order
__device__ __forceinline__ bool lattice_any_generic(const Lattice& l, const NarrowJob& job, const uint32_t* sm,
                                                    int lane) {
  const int npts = l.nu * l.nv;
  uint32_t hit = 0;
  int a = lane / l.nv, b = lane - a * l.nv;
  const int da = 32 / l.nv, db = 32 - da * l.nv;
  for (int p = lane; p < npts; p += 32) {
    const uint32_t fi = static_cast<uint32_t>(job.oi + a * job.ui + b * job.vi);
    const uint32_t fj = static_cast<uint32_t>(job.oj + a * job.uj + b * job.vj);
    const uint32_t w = sm[(fj >> 29) * 96 + (fi >> kFracBits)];
    hit |= w >> ((fj >> kFracBits) & 31u);
    a += da; b += db;
    if (b >= l.nv) { b -= l.nv; a += 1; }
  }
  return __any_sync(0xffffffffu, (hit & 1u) != 0);
}

// Fused broad + narrow: each warp takes 32 consecutive poses; the lanes set them up in
// parallel (fixed-point lattice, bounding box, tile / 8x8-block occupancy), then the warp
// walks the survivors: stage window -> exact box test -> lattice.
//   NU == 0 selects the run-time lattice shape.
template <int NU, int NV, int S>
__global__ void __launch_bounds__(kThreads)
collide_oriented_kernel(GridDev g, Lattice lat, const float* __restrict__ poses, uint8_t* __restrict__ flags,
                        int64_t n, int broad_phase, unsigned long long* __restrict__ narrow_counter) {
  __shared__ uint32_t smem[kWarpsPerBlock][kWinWords];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = blockIdx.x * static_cast<int64_t>(kWarpsPerBlock) + wib;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  const int64_t n_groups = (n + 31) >> 5;
  uint32_t* sm = smem[wib];
  unsigned long long local_narrow = 0;
  const float inv_res_f = static_cast<float>(1.0 / g.res);
  // half diagonal of the footprint in cells, + 1 cell for the floor, + 2 cells of float slack
  const int reach = static_cast<int>(sqrt(lat.su * lat.hu * lat.su * lat.hu + lat.sv * lat.hv * lat.sv * lat.hv)) + 4;
  for (int64_t grp = warp; grp < n_groups; grp += n_warps) {
    const int64_t k = grp * 32 + lane;
    NarrowJob job;
    bool need = false;
    if (k < n) {
      const float x = poses[3 * k], y = poses[3 * k + 1], yaw = poses[3 * k + 2];
      // cheapest test first: no 8x8 block with an obstacle inside the square that bounds the
      // footprint for ANY yaw (float arithmetic + 2 cells of slack: conservative, never decisive
      // for a colliding pose).  Only survivors pay for the fp64 lattice set-up.
      bool maybe = true;
      if (broad_phase) {
        const float cif = static_cast<float>(g.W / 2) - y * inv_res_f;
        const float cjf = static_cast<float>(g.H / 2) - x * inv_res_f;
        if (fabsf(cif) < 1.0e6f && fabsf(cjf) < 1.0e6f) {
          const int ic = __float2int_rd(cif), jc = __float2int_rd(cjf);
          const int i0 = ic - reach, i1 = ic + reach, j0 = jc - reach, j1 = jc + reach;
          if (i1 < 0 || i0 >= g.W || j1 < 0 || j0 >= g.H) maybe = false;
          else maybe = coarse_square_any(g, max(i0, -kHaloCells), min(i1, g.W + kHaloCells - 1),
                                         max(j0, -kHaloCells), min(j1, g.H + kHaloCells - 1));
        }
      }
      if (maybe) {
        const PoseFx p = pose_to_fixed(lat, g.W, g.H, g.res, static_cast<double>(x), static_cast<double>(y),
                                       static_cast<double>(yaw));
        need = make_job(g, lat, p, broad_phase, &job);
      }
    }
    uint32_t todo = __ballot_sync(0xffffffffu, need);
    local_narrow += __popc(todo);
    bool my_hit = false;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      NarrowJob j;
      j.oi = __shfl_sync(0xffffffffu, job.oi, src);
      j.oj = __shfl_sync(0xffffffffu, job.oj, src);
      j.ui = __shfl_sync(0xffffffffu, job.ui, src);
      j.uj = __shfl_sync(0xffffffffu, job.uj, src);
      j.vi = __shfl_sync(0xffffffffu, job.vi, src);
      j.vj = __shfl_sync(0xffffffffu, job.vj, src);
      j.tiles = __shfl_sync(0xffffffffu, job.tiles, src);
      j.box = __shfl_sync(0xffffffffu, job.box, src);
      j.occ = __shfl_sync(0xffffffffu, job.occ, src);
      stage_window(g, j, sm, lane);
      bool hit = false;
      if (!broad_phase || window_box_any(j, sm, lane)) {
        if (NU > 0) hit = lattice_any_fixed<(NU > 0 ? NU : 1), (NU > 0 ? NV : 1), S>(j, sm, lane);
        else hit = lattice_any_generic(lat, j, sm, lane);
      }
      __syncwarp();
      if (lane == src) my_hit = hit;
    }
    if (k < n) flags[k] = my_hit ? 1 : 0;
  }
  if (lane == 0 && narrow_counter && local_narrow) atomicAdd(narrow_counter, local_narrow);
}

__global__ void __launch_bounds__(kThreads)
collide_edges_kernel(GridDev g, CollideParams cp, const float* __restrict__ edges, uint8_t* __restrict__ flags,
                     int64_t n) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* e = edges + 5 * k;
    flags[k] = collide_edge_thread(g, cp, static_cast<double>(e[0]), static_cast<double>(e[1]),
                                   static_cast<double>(e[2]), static_cast<double>(e[3]),
                                   static_cast<double>(e[4])) ? 1 : 0;
  }
}

__global__ void footprint_points_kernel(GridDev g, Lattice l, double x, double y, double yaw, double* out,
                                        int cap, int* n_out) {
  const PoseFx p = pose_to_fixed(l, g.W, g.H, g.res, x, y, yaw);
  if (!p.valid) { if (threadIdx.x == 0) *n_out = 0; return; }
  const int npts = l.nu * l.nv;
  for (int k = threadIdx.x; k < npts && k < cap; k += blockDim.x) {
    const int a = k / l.nv, b = k - a * l.nv;
    const int32_t fi = p.oi + a * p.ui + b * p.vi;
    const int32_t fj = p.oj + a * p.uj + b * p.vj;
    const double ci = dm::add(static_cast<double>(p.ai), dm::div(static_cast<double>(fi), kFixedOne));
    const double cj = dm::add(static_cast<double>(p.aj), dm::div(static_cast<double>(fj), kFixedOne));
    out[2 * k + 0] = dm::mul(dm::sub(static_cast<double>(g.H / 2), cj), g.res);
    out[2 * k + 1] = dm::mul(dm::sub(static_cast<double>(g.W / 2), ci), g.res);
  }
  if (threadIdx.x == 0) *n_out = npts < cap ? npts : cap;
}

int grid_blocks(const pp_handle* h, int64_t work_items, int per_block, int blocks_per_sm) {
  const int64_t want = (work_items + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(h->sm_count) * blocks_per_sm;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

// one launch over n device-resident poses; the caller owns counter reset / read-back
int collide_launch(pp_handle* h, int64_t n, const float* poses, uint8_t* flags) {
  if (n == 0) return PP_OK;
  cudaStream_t s = h->stream;
  const int mode = h->P.collision_mode;
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(h->d_count);
  if (mode == PP_COLLISION_AS_SHIPPED) {
    fill_zero_kernel<<<grid_blocks(h, n, kThreads * 16, 8), kThreads, 0, s>>>(flags, n);
  } else if (mode == PP_COLLISION_REF_AABB) {
    collide_aabb_kernel<<<grid_blocks(h, n, kThreads, 8), kThreads, 0, s>>>(h->grid, poses, flags, n);
  } else {
    const int blocks = grid_blocks(h, (n + 31) / 32, kWarpsPerBlock, 8);
    if (h->lat.nu == 48 && h->lat.nv == 17)        // 4.7 m x 1.586 m at 0.1 m cells
      collide_oriented_kernel<48, 17, 2><<<blocks, kThreads, 0, s>>>(h->grid, h->lat, poses, flags, n, h->broad_phase, counter);
    else if (h->lat.nu == 20 && h->lat.nv == 8)    // the same car at the launch file's 0.25 m cells
      collide_oriented_kernel<20, 8, 4><<<blocks, kThreads, 0, s>>>(h->grid, h->lat, poses, flags, n, h->broad_phase, counter);
    else
      collide_oriented_kernel<0, 0, 1><<<blocks, kThreads, 0, s>>>(h->grid, h->lat, poses, flags, n, h->broad_phase, counter);
  }
  h->launches++;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}

int collide_counter_reset(pp_handle* h) {
  PP_CUDA(cudaMemsetAsync(h->d_count, 0, sizeof(unsigned long long), h->stream));
  return PP_OK;
}
int collide_counter_fetch(pp_handle* h) {
  PP_CUDA(cudaMemcpyAsync(h->h_count, h->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  h->last_narrow = -1;  // resolved from the pinned mirror after the next sync
  return PP_OK;
}

int collide_batch_dev(pp_handle* h, int64_t n, const float* poses, uint8_t* flags) {
  if (n == 0) return PP_OK;
  cudaStream_t s = h->stream;
  int rc = collide_counter_reset(h);
  if (rc) return rc;
  PP_CUDA(cudaEventRecord(h->ev[2], s));
  rc = collide_launch(h, n, poses, flags);
  if (rc) return rc;
  PP_CUDA(cudaEventRecord(h->ev[3], s));
  rc = collide_counter_fetch(h);
  if (rc) return rc;
  if (h->P.collision_mode != PP_COLLISION_ORIENTED) h->last_narrow = n;
  return PP_OK;
}

int collide_edges_dev(pp_handle* h, int64_t n, const float* edges, uint8_t* flags) {
  if (n == 0) return PP_OK;
  collide_edges_kernel<<<grid_blocks(h, n, kThreads, 8), kThreads, 0, h->stream>>>(h->grid, collide_params(h),
                                                                                  edges, flags, n);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}

int footprint_points_dev(pp_handle* h, double x, double y, double yaw, double* out_dev, int cap, int* n_dev) {
  footprint_points_kernel<<<1, 256, 0, h->stream>>>(h->grid, h->lat, x, y, yaw, out_dev, cap, n_dev);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}

}  // namespace pp
