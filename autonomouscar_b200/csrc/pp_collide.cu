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
#include "pp_narrow.cuh"

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
