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

// What the narrow phase needs to know about one pose, relative to the first tile under
// its bounding box (so all lattice coordinates are in [0, 96) cells).
struct NarrowJob {
  int32_t oi, oj, ui, uj, vi, vj;
  int t0i, t0j;   // first tile row / column (tile units)
  int nti, ntj;   // tiles under the box, 1..3 each
};

__device__ __forceinline__ bool make_job(const GridDev& g, const Lattice& l, const PoseFx& p, NarrowJob* job) {
  if (!p.valid) return false;
  int i0, i1, j0, j1;
  lattice_bbox(l, p, &i0, &i1, &j0, &j1);
  if (i1 < 0 || i0 >= g.W || j1 < 0 || j0 >= g.H) return false;  // box entirely off the grid
  job->t0i = (i0 + kHaloCells) >> 5;
  job->t0j = (j0 + kHaloCells) >> 5;
  job->nti = ((i1 + kHaloCells) >> 5) - job->t0i + 1;
  job->ntj = ((j1 + kHaloCells) >> 5) - job->t0j + 1;
  // re-base the lattice origin from the anchor cell to the first tile's corner
  const int shift_i = p.ai - (job->t0i * kTile - kHaloCells);
  const int shift_j = p.aj - (job->t0j * kTile - kHaloCells);
  job->oi = p.oi + (shift_i << kFracBits);
  job->oj = p.oj + (shift_j << kFracBits);
  job->ui = p.ui; job->uj = p.uj; job->vi = p.vi; job->vj = p.vj;
  return true;
}

__device__ __forceinline__ bool tiles_have_obstacle(const GridDev& g, const NarrowJob& job) {
  uint32_t any = 0;
  for (int a = 0; a < job.nti; ++a)
    for (int b = 0; b < job.ntj; ++b)
      any |= __ldg(&g.tile_any[static_cast<size_t>(job.t0i + a) * g.TJ + (job.t0j + b)]);
  return any != 0;
}

// The warp evaluates the lattice of ONE pose.  smem: 9 tiles x 32 words, private to the warp.
__device__ __forceinline__ bool narrow_phase_warp(const GridDev& g, const Lattice& l, const NarrowJob& job,
                                                  uint32_t* smem_tiles, int lane) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      uint32_t w = 0;
      if (a < job.nti && b < job.ntj)
        w = __ldg(&g.occ_tiles[(static_cast<size_t>(job.t0i + a) * g.TJ + (job.t0j + b)) * kTile + lane]);
      smem_tiles[(a * 3 + b) * kTile + lane] = w;
    }
  }
  __syncwarp();
  const int npts = l.nu * l.nv;
  uint32_t hit = 0;
  // lane-strided walk over the lattice in (a, b) order; (a, b) advanced incrementally
  int a = lane / l.nv, b = lane - a * l.nv;
  const int da = 32 / l.nv, db = 32 - da * l.nv;
  for (int p = lane; p < npts; p += 32) {
    const int32_t fi = job.oi + a * job.ui + b * job.vi;
    const int32_t fj = job.oj + a * job.uj + b * job.vj;
    const int ri = fi >> kFracBits, rj = fj >> kFracBits;
    const uint32_t w = smem_tiles[((ri >> 5) * 3 + (rj >> 5)) * kTile + (ri & 31)];
    hit |= w >> (rj & 31);
    a += da; b += db;
    if (b >= l.nv) { b -= l.nv; a += 1; }
  }
  const bool any = __any_sync(0xffffffffu, (hit & 1u) != 0);
  __syncwarp();
  return any;
}

// Fused broad + narrow: each warp takes 32 consecutive poses; lanes set them up in
// parallel, then the warp walks the ones that need the narrow phase.
__global__ void __launch_bounds__(kThreads)
collide_oriented_kernel(GridDev g, Lattice lat, const float* __restrict__ poses, uint8_t* __restrict__ flags,
                        int64_t n, int broad_phase, unsigned long long* __restrict__ narrow_counter) {
  __shared__ uint32_t smem[kWarpsPerBlock][9 * kTile];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = blockIdx.x * static_cast<int64_t>(kWarpsPerBlock) + wib;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  const int64_t n_groups = (n + 31) >> 5;
  unsigned long long local_narrow = 0;
  for (int64_t grp = warp; grp < n_groups; grp += n_warps) {
    const int64_t k = grp * 32 + lane;
    NarrowJob job;
    bool need = false;
    if (k < n) {
      const float x = poses[3 * k], y = poses[3 * k + 1], yaw = poses[3 * k + 2];
      const PoseFx p = pose_to_fixed(lat, g.W, g.H, g.res, static_cast<double>(x), static_cast<double>(y),
                                     static_cast<double>(yaw));
      need = make_job(g, lat, p, &job);
      if (need && broad_phase) need = tiles_have_obstacle(g, job);
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
      j.t0i = __shfl_sync(0xffffffffu, job.t0i, src);
      j.t0j = __shfl_sync(0xffffffffu, job.t0j, src);
      j.nti = __shfl_sync(0xffffffffu, job.nti, src);
      j.ntj = __shfl_sync(0xffffffffu, job.ntj, src);
      const bool hit = narrow_phase_warp(g, lat, j, smem[wib], lane);
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
    collide_oriented_kernel<<<grid_blocks(h, (n + 31) / 32, kWarpsPerBlock, 8), kThreads, 0, s>>>(
        h->grid, h->lat, poses, flags, n, h->broad_phase, counter);
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
