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
#include <cstdlib>

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

// One launch resolves a whole batch.  A CTA takes 1024 consecutive poses at a time and runs
// three phases, each with dense lanes (survivors are compacted through shared memory):
//   1. thread per pose : one bit of the dilated 8x8 map ("any obstacle within reach of this
//                        block, for any yaw") decides ~70 % of the poses; their flag is 0
//   2. thread per survivor : fp64 fixed-point lattice set-up, bounding box, tile / 8x8-block
//                        occupancy under the box -> NarrowJob or flag 0
//   3. warp per job    : stage the occupied tiles under the box (one 128 B line each), exact
//                        box test on the staged window, then the sample lattice
//   NU == 0 selects the run-time lattice shape.  broad_phase == 0 ("direct") disables every
//   pruning test: each pose that overlaps the grid runs the full lattice.
#ifndef PP_COLLIDE_PPT
#define PP_COLLIDE_PPT 4
#endif
constexpr int kPosesPerThread = PP_COLLIDE_PPT;
constexpr int kChunk = kThreads * kPosesPerThread;

__device__ __forceinline__ int warp_append(int* counter, bool pred, int lane) {
  const unsigned bal = __ballot_sync(0xffffffffu, pred);
  int base = 0;
  if (lane == 0 && bal) base = atomicAdd(counter, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + __popc(bal & ((1u << lane) - 1u));
}

#ifndef PP_LATTICE_S
#define PP_LATTICE_S 2
#endif
#ifndef PP_COLLIDE_MIN_BLOCKS
#define PP_COLLIDE_MIN_BLOCKS 4
#endif
template <int NU, int NV, int S>
__global__ void __launch_bounds__(kThreads, PP_COLLIDE_MIN_BLOCKS)
collide_oriented_kernel(GridDev g, Lattice lat, const float* __restrict__ poses, uint8_t* __restrict__ flags,
                        int64_t n, int broad_phase, unsigned long long* __restrict__ narrow_counter) {
  __shared__ __align__(16) uint32_t s_win[kWarpsPerBlock][kWinWords];
  __shared__ int s_queue[kChunk];
  __shared__ NarrowJob s_jobs[kThreads];
  __shared__ int s_job_off[kThreads];
  __shared__ int s_nq, s_nj;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const float inv_res_f = static_cast<float>(1.0 / g.res);
  const int reach = g.reach_cells;
  unsigned long long local_narrow = 0;
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t base = chunk * kChunk;
    if (tid == 0) s_nq = 0;
    __syncthreads();
    // ---- phase 1 ------------------------------------------------------------------------
#pragma unroll
    for (int p = 0; p < kPosesPerThread; ++p) {
      const int off = p * kThreads + tid;
      const int64_t k = base + off;
      bool survive = false;
      if (k < n) {
        survive = true;
        if (broad_phase) {
          const float x = poses[3 * k], y = poses[3 * k + 1];
          const float cif = static_cast<float>(g.W / 2) - y * inv_res_f;
          const float cjf = static_cast<float>(g.H / 2) - x * inv_res_f;
          if (fabsf(cif) < 1.0e6f && fabsf(cjf) < 1.0e6f) {   // (NaN / absurd poses take the exact path)
            const int ic = __float2int_rd(cif), jc = __float2int_rd(cjf);
            if (ic + reach < 0 || ic - reach >= g.W || jc + reach < 0 || jc - reach >= g.H) {
              survive = false;                               // the footprint cannot reach the grid
            } else {
              const int r8 = (ic + kHaloCells) >> 3, c8 = (jc + kHaloCells) >> 3;
              survive = (__ldg(&g.near8[static_cast<size_t>(r8) * g.CW8 + (c8 >> 6)]) >> (c8 & 63)) & 1ull;
            }
          }
        }
        if (!survive) flags[k] = 0;
      }
      const int slot = warp_append(&s_nq, survive, lane);
      if (survive) s_queue[slot] = off;
    }
    __syncthreads();
    const int nq = s_nq;
    // ---- phases 2 + 3, 256 survivors at a time ------------------------------------------------
    for (int q0 = 0; q0 < nq; q0 += kThreads) {
      if (tid == 0) s_nj = 0;
      __syncthreads();
      const int qi = q0 + tid;
      NarrowJob job;
      bool need = false;
      int off = 0;
      if (qi < nq) {
        off = s_queue[qi];
        const int64_t k = base + off;
        const float x = poses[3 * k], y = poses[3 * k + 1], yaw = poses[3 * k + 2];
        const PoseFx px = pose_to_fixed(lat, g.W, g.H, g.res, static_cast<double>(x), static_cast<double>(y),
                                        static_cast<double>(yaw));
        need = make_job(g, lat, px, broad_phase, &job);
        if (!need) flags[k] = 0;
      }
      const int slot = warp_append(&s_nj, need, lane);
      if (need) { s_jobs[slot] = job; s_job_off[slot] = off; }
      __syncthreads();
      const int nj = s_nj;
      if (tid == 0) local_narrow += nj;
      for (int j = wib; j < nj; j += kWarpsPerBlock) {
        const NarrowJob jb = s_jobs[j];
        stage_window(g, jb, s_win[wib], lane);
        bool hit = false;
        if (!broad_phase || window_box_any(jb, s_win[wib], lane)) {
          if (NU > 0) hit = lattice_any_fixed<(NU > 0 ? NU : 1), (NU > 0 ? NV : 1), S>(jb, s_win[wib], lane);
          else hit = lattice_any_generic(lat, jb, s_win[wib], lane);
        }
        __syncwarp();
        if (lane == 0) flags[base + s_job_off[j]] = hit ? 1 : 0;
      }
      __syncthreads();
    }
  }
  if (tid == 0 && narrow_counter && local_narrow) atomicAdd(narrow_counter, local_narrow);
}

// Along-edge test (row S3), thread per edge: every step evaluated by one thread.  Kept for the
// non-ORIENTED modes (one map lookup per step) and as the small-batch path.
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

// ORIENTED along-edge test, warp per edge: the steps of the edge are a mini-batch -- lane l sets up
// step base + l + 1 (fixed-point lattice, broad phase: one thread each), then the warp runs the
// staged-window narrow phase only for the steps that need it, stopping at the first hit.  Same
// step rule as collide_edge_thread / the oracle: n = max(int(dist / search_res * 2), 1) poses at
// t = k / n, k = 1..n (spacing rule of LP:308).
__global__ void __launch_bounds__(kThreads)
collide_edges_warp_kernel(GridDev g, CollideParams cp, const float* __restrict__ edges, uint8_t* __restrict__ flags,
                          int64_t n) {
  __shared__ __align__(16) uint32_t s_win[kWarpsPerBlock][kWinWords];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t e = warp; e < n; e += n_warps) {
    const float* ed = edges + 5 * e;
    const double x0 = ed[0], y0 = ed[1], x1 = ed[2], y1 = ed[3], yaw = ed[4];
    const double dx = dm::sub(x1, x0), dy = dm::sub(y1, y0);
    const double dist = __dsqrt_rn(dm::add(dm::mul(dx, dx), dm::mul(dy, dy)));
    int steps;
    bool hit = false;
    if (dm::trunc_to_int(dm::mul(dm::div(dist, cp.search_res), 2.0), &steps)) {
      if (steps < 1) steps = 1;
      for (int base = 0; base < steps && !hit; base += 32) {
        const int k = base + lane + 1;
        bool need = false;
        NarrowJob job;
        if (k <= steps) {
          const double t = dm::div(static_cast<double>(k), static_cast<double>(steps));
          const PoseFx pf = pose_to_fixed(cp.lat, g.W, g.H, g.res, dm::add(x0, dm::mul(t, dx)), dm::add(y0, dm::mul(t, dy)), yaw);
          need = make_job(g, cp.lat, pf, 1, &job);
        }
        hit = warp_collide_jobs(g, cp.lat, job, need, s_win[wib], lane);
      }
    }
    if (lane == 0) flags[e] = hit ? 1 : 0;
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
    static const int bps = getenv("PP_COLLIDE_BLOCKS_PER_SM") ? atoi(getenv("PP_COLLIDE_BLOCKS_PER_SM")) : 64;
    const int blocks = grid_blocks(h, n, kChunk, bps);
    if (h->lat.nu == 48 && h->lat.nv == 17)        // 4.7 m x 1.586 m at 0.1 m cells
      collide_oriented_kernel<48, 17, PP_LATTICE_S><<<blocks, kThreads, 0, s>>>(h->grid, h->lat, poses, flags, n, h->broad_phase, counter);
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
  if (h->P.collision_mode == PP_COLLISION_ORIENTED && !getenv("PP_EDGES_THREAD_KERNEL"))
    collide_edges_warp_kernel<<<grid_blocks(h, n, kWarpsPerBlock, 16), kThreads, 0, h->stream>>>(h->grid, collide_params(h),
                                                                                                edges, flags, n);
  else
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
