// pp_rrt.cu -- the sampling planner BASELINE.json:north_star names (rows S1-S3 fused):
// Philox sample -> nearest neighbour over the tree -> steer with the reference's motion
// model (local_planner.cpp:155-157) -> oriented-footprint collision along the edge
// (spacing rule of local_planner.cpp:308) -> append.
//
// The reference planner has no sampling loop; the semantics are defined by
// this build (DESIGN.md "Sampling planner") and restated on the CPU in
// oracle/sampling_ref.h.  The loop is ROUND-SYNCHRONOUS so that the result does not depend
// on the parallel schedule: round r resolves samples r*K .. r*K+K-1 against the tree as of
// the end of round r-1 and appends the survivors in sample order.
//
// Three kernels per round, for all queries of a batch at once:
//   rrt_nearest  grid (node slices, queries): every CTA scans its slice of the tree (staged
//                through shared memory) for all K samples of the round; partial results are
//                merged with one 64-bit atomicMin per (sample, slice) on (dist2 bits, index),
//                which is exactly "minimum distance, ties -> lowest index"
//   rrt_extend   one warp per (query, sample): steer, then the warp-cooperative footprint
//                test (pp_narrow.cuh) at every edge step
//   rrt_append   one CTA per query: order-preserving compaction of the survivors, goal test,
//                termination
// Samples are regenerated from Philox(seed, query_id, sample index) wherever needed and are
// never written to memory.
#include <algorithm>
#include <cstring>
#include <vector>

#include "pp_common.cuh"
#include "pp_narrow.cuh"
#include "pp_philox.cuh"

namespace pp {

namespace {

constexpr int kT = 256;
constexpr int kMaxK = 1024;               // samples per round the kernels support
constexpr int kNodeTile = 1024;           // float2 nodes staged per step
constexpr int kSliceGrain = 128;          // node slices handed to CTAs are multiples of this
constexpr int kExtSplitMax = 4;           // warps that share the poses of one edge in rrt_extend (few samples per round)
constexpr unsigned long long kBestInit = ~0ull;

struct RrtState {
  int n_nodes, n_samples, n_rounds, n_coll, goal_id, done;
};
struct RrtHdr {
  int status, n_rounds, n_nodes, n_samples, n_coll, n_path;
  unsigned long long digest;
};

struct RrtArgs {
  GridDev g;
  CollideParams cp;
  pp_params P;
  const pp_query* queries;
  int n_queries;
  int K;               // samples per round
  int NC;              // node capacity per query
  int cap_path;
  float x_lo, x_rng, y_lo, y_rng, bias;
  // per query SoA node storage: [q][NC]
  double *nx, *ny, *nhd, *nvel, *nstep;
  int* nparent;
  float2* nxy;
  // per round scratch: [q][K]
  unsigned long long* best;
  double* cand;        // [q][K][5] : x y yaw v step
  int* cand_near;      // [q][K]
  int* cand_valid;     // [q][K]  1 = append, 0 = dropped, 2 = dropped by collision
  RrtState* state;
  RrtHdr* hdr;
  double* out_path;    // [q][cap_path][2]
  int* out_path_ids;   // [q][cap_path]
  double* out_profile; // [q][3][cap_path]
  int* n_active;       // number of queries not yet done
};

__device__ __forceinline__ void draw_sample(const RrtArgs& A, const pp_query& q, long long s, float* sx, float* sy) {
  const Philox4 r = philox4x32_10(static_cast<uint32_t>(s), static_cast<uint32_t>(static_cast<unsigned long long>(s) >> 32),
                                  q.query_id, kTagRrt, static_cast<uint32_t>(q.seed),
                                  static_cast<uint32_t>(q.seed >> 32));
  if (u01_from_bits(r.z) < A.bias) {
    *sx = static_cast<float>(q.goal_x);
    *sy = static_cast<float>(q.goal_y);
  } else {
    *sx = __fmaf_rn(u01_from_bits(r.x), A.x_rng, A.x_lo);
    *sy = __fmaf_rn(u01_from_bits(r.y), A.y_rng, A.y_lo);
  }
}

__global__ void __launch_bounds__(kT) rrt_init(RrtArgs A) {
  const int q = blockIdx.x;
  if (q >= A.n_queries) return;
  const pp_query qq = A.queries[q];
  const size_t nb = static_cast<size_t>(q) * A.NC;
  if (threadIdx.x == 0) {
    A.nx[nb] = qq.start_x; A.ny[nb] = qq.start_y; A.nhd[nb] = qq.start_heading; A.nvel[nb] = qq.start_speed;
    A.nstep[nb] = 0.0; A.nparent[nb] = -1;
    A.nxy[nb] = make_float2(static_cast<float>(qq.start_x), static_cast<float>(qq.start_y));
    RrtState st;
    st.n_nodes = 1; st.n_samples = 0; st.n_rounds = 0; st.n_coll = 0; st.goal_id = -1;
    // the oracle's loop condition, evaluated before the first round
    st.done = !(0 < A.P.rrt_max_samples && 1 < A.P.max_nodes);
    A.state[q] = st;
    if (!st.done) atomicAdd(A.n_active, 1);
  }
  for (int k = threadIdx.x; k < A.K; k += kT) {
    A.best[static_cast<size_t>(q) * A.K + k] = kBestInit;
    A.cand_valid[static_cast<size_t>(q) * A.K + k] = 0;
  }
}

// SPT = samples owned by one thread (ceil(K / 256), 1 / 2 / 4): the distance loop is unrolled over exactly
// the samples that exist.
template <int SPT>
__global__ void __launch_bounds__(kT) rrt_nearest(RrtArgs A, int slice_nodes) {
  __shared__ float2 tile[kNodeTile];
  const int q = blockIdx.y;
  const RrtState st = A.state[q];
  if (st.done) return;
  const int n_nodes = st.n_nodes;
  const int lo = blockIdx.x * slice_nodes;
  if (lo >= n_nodes) return;
  const int hi = min(lo + slice_nodes, n_nodes);
  const pp_query qq = A.queries[q];
  const long long s0 = st.n_samples;
  const int k_round = static_cast<int>(min(static_cast<long long>(A.K), static_cast<long long>(A.P.rrt_max_samples) - s0));
  // K >= 256: every thread owns up to SPT samples and scans the whole tile.
  // K <  256: the threads form G = 256 / Kp groups (Kp = K rounded up to a power of two); a
  // group's thread owns one sample and scans every G-th node of the tile, so no lane idles.
  int Kp = 1;
  while (Kp < A.K) Kp <<= 1;
  const int G = Kp >= kT ? 1 : kT / Kp;
  const int grp = G == 1 ? 0 : threadIdx.x / Kp;
  float sx[SPT], sy[SPT], bd[SPT];
  int bi[SPT], st_[SPT];
#pragma unroll
  for (int u = 0; u < SPT; ++u) {
    const int t = G == 1 ? threadIdx.x + u * kT : (u == 0 ? threadIdx.x % Kp : k_round);
    st_[u] = t;
    sx[u] = 0.f; sy[u] = 0.f; bd[u] = 0.f; bi[u] = -1;
    if (t < k_round) draw_sample(A, qq, s0 + t, &sx[u], &sy[u]);
  }
  const float2* nodes = A.nxy + static_cast<size_t>(q) * A.NC;
  for (int base = lo; base < hi; base += kNodeTile) {
    const int cnt = min(kNodeTile, hi - base);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += kT) tile[k] = nodes[base + k];
    __syncthreads();
    if (G == 1) {
      for (int k = 0; k < cnt; ++k) {
        const float2 nd = tile[k];
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
          const float dx = __fsub_rn(sx[u], nd.x), dy = __fsub_rn(sy[u], nd.y);
          const float d = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
          if (bi[u] < 0 || d < bd[u]) { bd[u] = d; bi[u] = base + k; }
        }
      }
    } else {
      for (int k = grp; k < cnt; k += G) {
        const float2 nd = tile[k];
        const float dx = __fsub_rn(sx[0], nd.x), dy = __fsub_rn(sy[0], nd.y);
        const float d = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
        if (bi[0] < 0 || d < bd[0]) { bd[0] = d; bi[0] = base + k; }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < SPT; ++u) {
    const int t = st_[u];
    if (t < k_round && bi[u] >= 0) {
      const unsigned long long packed = (static_cast<unsigned long long>(__float_as_uint(bd[u])) << 32) |
                                        static_cast<unsigned int>(bi[u]);
      atomicMin(&A.best[static_cast<size_t>(q) * A.K + t], packed);
    }
  }
}

// steer from tree node `near` toward the sample (sx, sy): heading change clamped to +-heading_diff, speed
// v' = min(v + a dt, max_speed), step (v + v') dt (LP:155-157).  Shared by the per-round kernels and the
// persistent kernel, so both build the same tree bit for bit.
struct Steer {
  double px, py, yaw, v_next, step, nx, ny, dx, dy;
  int n;        // edge poses k = 1..n (LP:308 spacing), 0 = no edge (zero step / unrepresentable length)
  bool moves;   // step > 0
};
__device__ __forceinline__ Steer steer_from(const RrtArgs& A, const pp_query& qq, int near, float sx, float sy) {
  Steer S;
  const size_t nb = near;
  S.px = __ldcg(&A.nx[nb]); S.py = __ldcg(&A.ny[nb]);
  const double ph = __ldcg(&A.nhd[nb]), pv = __ldcg(&A.nvel[nb]);
  const double th = dm::atan2_(dm::sub(static_cast<double>(sy), S.py), dm::sub(static_cast<double>(sx), S.px));
  double d = dm::wrap_pi(dm::sub(th, ph));
  if (d > A.P.heading_diff) d = A.P.heading_diff;
  if (d < -A.P.heading_diff) d = -A.P.heading_diff;
  S.yaw = dm::add(ph, d);
  S.v_next = dm::add(pv, dm::div(dm::mul(qq.max_accel, A.P.time_step_ms), 1000.0));
  if (S.v_next > qq.max_speed) S.v_next = qq.max_speed;
  const double sum_v = dm::add(pv, S.v_next);                                  // LP:155
  S.step = dm::div(dm::mul(sum_v, A.P.time_step_ms), 1000.0);
  S.moves = S.step > 0;
  S.nx = 0; S.ny = 0; S.dx = 0; S.dy = 0; S.n = 0;
  if (S.moves) {
    double sn, cs;
    dm::sincos(S.yaw, &sn, &cs);
    S.nx = dm::add(S.px, dm::mul(S.step, cs));                                  // LP:156
    S.ny = dm::add(S.py, dm::mul(S.step, sn));                                  // LP:157
    // edge collision, poses k = 1..max(n,1) with n = int(dist / search_res * 2) (LP:308)
    S.dx = dm::sub(S.nx, S.px); S.dy = dm::sub(S.ny, S.py);
    const double dist = __dsqrt_rn(dm::add(dm::mul(S.dx, S.dx), dm::mul(S.dy, S.dy)));
    int n = 0;
    if (dm::trunc_to_int(dm::mul(dm::div(dist, A.P.search_res), 2.0), &n)) S.n = n < 1 ? 1 : n;
  }
  return S;
}
// poses k_lo+1 .. k_hi of the edge, by one warp: the poses are a mini-batch -- lane l sets up pose base + l + 1
// (broad phase, one thread each), then the warp runs the narrow phase only for the poses that need it
__device__ __forceinline__ bool edge_part_hit(const RrtArgs& A, const Steer& S, int k_lo, int k_hi, uint32_t* smem, int lane) {
  bool hit = false;
  for (int base = k_lo; base < k_hi && !hit; base += 32) {
    const int k = base + lane + 1;
    bool need = false;
    NarrowJob job;
    if (k <= k_hi) {
      const double tt = dm::div(static_cast<double>(k), static_cast<double>(S.n));
      const double ex = dm::add(S.px, dm::mul(tt, S.dx)), ey = dm::add(S.py, dm::mul(tt, S.dy));
      if (A.cp.mode == PP_COLLISION_ORIENTED) {
        const PoseFx pf = pose_to_fixed(A.cp.lat, A.g.W, A.g.H, A.g.res, ex, ey, S.yaw);
        need = make_job_light(A.g, A.cp.lat, pf, &job);
      } else if (A.cp.mode == PP_COLLISION_REF_AABB) {
        hit = aabb_hit(A.g, ex, ey);
      }
    }
    if (A.cp.mode == PP_COLLISION_ORIENTED) hit = warp_collide_jobs(A.g, A.cp.lat, job, need, smem, lane);
    else hit = __any_sync(0xffffffffu, hit);
  }
  return hit;
}

__global__ void __launch_bounds__(kT) rrt_extend(RrtArgs A, int ext_split) {
  __shared__ __align__(16) uint32_t smem[kT / 32][kWinWords];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // ext_split warps share one sample: all of them steer (identical arithmetic), each tests its share of the
  // edge's poses; verdicts are combined with atomicMax (2 = collision beats 1 = free); cand_valid is zero at
  // the start of every round (rrt_init / rrt_append)
  const long long gwarp = blockIdx.x * static_cast<long long>(kT / 32) + wib;
  const int part = static_cast<int>(gwarp % ext_split);
  const long long warp = gwarp / ext_split;
  const int q = static_cast<int>(warp / A.K), t = static_cast<int>(warp % A.K);
  if (q >= A.n_queries) return;
  const RrtState st = A.state[q];
  if (st.done) return;
  const long long s = static_cast<long long>(st.n_samples) + t;
  const size_t slot = static_cast<size_t>(q) * A.K + t;
  if (s >= A.P.rrt_max_samples) return;
  const pp_query qq = A.queries[q];
  float sx, sy;
  draw_sample(A, qq, s, &sx, &sy);
  const int near = static_cast<int>(A.best[slot] & 0xffffffffull);
  RrtArgs Aq = A;                                    // node arrays of this query
  const size_t nb = static_cast<size_t>(q) * A.NC;
  Aq.nx += nb; Aq.ny += nb; Aq.nhd += nb; Aq.nvel += nb;
  const Steer S = steer_from(Aq, qq, near, sx, sy);
  int verdict = 0;
  if (S.moves) {
    bool hit = false;
    if (S.n > 0) {
      const int k_lo = static_cast<int>(static_cast<long long>(S.n) * part / ext_split);        // poses k_lo+1 .. k_hi
      const int k_hi = static_cast<int>(static_cast<long long>(S.n) * (part + 1) / ext_split);
      hit = edge_part_hit(A, S, k_lo, k_hi, smem[wib], lane);
    }
    verdict = hit ? 2 : 1;
  }
  if (lane == 0 && verdict) atomicMax(&A.cand_valid[slot], verdict);
  if (lane == 0 && part == 0) {
    A.cand_near[slot] = near;
    double* c = A.cand + slot * 5;
    c[0] = S.nx; c[1] = S.ny; c[2] = S.yaw; c[3] = S.v_next; c[4] = S.step;
  }
}

__global__ void __launch_bounds__(kT) rrt_append(RrtArgs A) {
  __shared__ int s_warp[kT / 32];
  __shared__ int s_goal;
  const int q = blockIdx.x;
  if (q >= A.n_queries) return;
  RrtState st = A.state[q];
  if (st.done) return;
  const pp_query qq = A.queries[q];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k_round = static_cast<int>(min(static_cast<long long>(A.K),
                                           static_cast<long long>(A.P.rrt_max_samples) - st.n_samples));
  if (threadIdx.x == 0) s_goal = 0x7fffffff;
  __syncthreads();
  int appended = 0, coll = 0;
  const size_t nb = static_cast<size_t>(q) * A.NC;
  for (int base = 0; base < k_round; base += kT) {
    const int t = base + threadIdx.x;
    const size_t slot = static_cast<size_t>(q) * A.K + t;
    const int verdict = t < k_round ? A.cand_valid[slot] : 0;
    if (t < k_round) A.cand_valid[slot] = 0;          // re-armed for the next round's atomicMax
    const bool ok = verdict == 1;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    coll += verdict == 2;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kT / 32; ++w) { const int v = s_warp[w]; if (w < warp) before += v; total += v; }
    if (ok) {
      const int id = st.n_nodes + appended + before + __popc(bal & ((1u << lane) - 1u));
      const double* c = A.cand + slot * 5;
      A.nx[nb + id] = c[0]; A.ny[nb + id] = c[1]; A.nhd[nb + id] = c[2]; A.nvel[nb + id] = c[3];
      A.nstep[nb + id] = c[4];
      A.nparent[nb + id] = A.cand_near[slot];
      A.nxy[nb + id] = make_float2(static_cast<float>(c[0]), static_cast<float>(c[1]));
      if (fabs(dm::sub(c[0], qq.goal_x)) <= A.P.rrt_goal_tolerance &&
          fabs(dm::sub(c[1], qq.goal_y)) <= A.P.rrt_goal_tolerance)
        atomicMin(&s_goal, id);   // the first appended node (lowest sample index) in the goal box
    }
    if (t < k_round) A.best[slot] = kBestInit;   // re-arm for the next round
    appended += total;
    __syncthreads();
  }
  // block-wide sum of collisions
  for (int off = 16; off > 0; off >>= 1) coll += __shfl_xor_sync(0xffffffffu, coll, off);
  if (lane == 0) s_warp[warp] = coll;
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < kT / 32; ++w) c += s_warp[w];
    st.n_nodes += appended;
    st.n_samples += k_round;
    st.n_rounds += 1;
    st.n_coll += c;
    if (s_goal != 0x7fffffff) st.goal_id = s_goal;
    st.done = !(st.goal_id < 0 && st.n_samples < A.P.rrt_max_samples && st.n_nodes < A.P.max_nodes);
    A.state[q] = st;
    if (st.done) atomicSub(A.n_active, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------
// One query, one kernel: the whole round loop runs inside a cooperative launch (all CTAs co-resident) with two
// grid-wide barriers per round instead of three launches, and the nearest-neighbour query goes through a bucket
// grid once the tree is large.  Results are those of the per-round kernels bit for bit (same sample stream, same
// float32 (distance, index) minimum, same steering / collision arithmetic, same append order).
//
//   phase A  CTA per sample: draw the sample, nearest node (exact, below), steer, edge collision with one warp per
//            share of the edge's poses -> verdict + candidate
//   ------ grid barrier ------
//   phase B  every CTA reads the K verdicts and derives the same survivor ranks (order-preserving append);
//            the CTA of a surviving sample writes its node, bins it, tests the goal box (atomicMin = first in
//            sample order).  The round counters are replicated in every CTA (a deterministic function of the
//            verdicts), so nothing central has to be waited for
//   ------ grid barrier ------
//   every ~16 rounds: the nodes are regrouped by bucket (counting sort: scan by CTA 0, scatter by all; two more
//   barriers), so that the next rounds scan whole buckets with coalesced loads
//
// Nearest neighbour (S2), exact: nodes [0, n_sorted) sit grouped by the cell of a kGridSide^2 bucket grid over
// the sampling box (16-byte records x, y, id), nodes [n_sorted, n_nodes) -- the "tail", <= kTailMax -- are scanned
// directly.  Every CTA keeps the cell offsets and the list of non-empty cells in shared memory.  Per sample:
// (1) the non-empty cell with the smallest box distance, its nodes (whole CTA) and the tail give an upper bound;
// (2) every other non-empty cell whose (conservatively rounded) box distance does not exceed it is scanned too,
//     one warp per cell so that the cells' L2 round trips overlap.
// The minimum over (dist2 bits << 32 | index) is exactly "minimum distance, ties -> lowest index".
#ifndef PP_RRT_WARP_CELLS
#define PP_RRT_WARP_CELLS 0
#endif
#ifndef PP_RRT_TAIL
#define PP_RRT_TAIL 4096
#endif
#ifndef PP_RRT_GRID
#define PP_RRT_GRID 48
#endif
constexpr int kGridSide = PP_RRT_GRID;             // (kGridSide^2 must be a multiple of the CTA size)
constexpr int kGridCells = kGridSide * kGridSide;
constexpr int kTailMax = PP_RRT_TAIL;
constexpr int kGridFrom = 8192;              // smaller trees are scanned whole
constexpr int kWorkCap = 1024;               // candidate cells of step (2) kept in shared memory

struct RrtPersist {
  int* cell_count;       // [cells]      nodes per cell, all nodes
  int* cell_start;       // [cells + 1]  first slot of each cell in the sorted arrays (nodes [0, n_sorted))
  int* cell_cursor;      // [cells]      scatter cursors of a rebuild
  float4* sorted;        // [NC]  (x, y, id as bits, 0) grouped by cell
  int* goal_id;          // atomicMin of the node ids inside the goal box (INT_MAX = none)
  unsigned* barrier;     // monotonic arrival counter of the grid barrier
  unsigned* tickets;     // [2] sample tickets of phase A, one counter per round parity (the other one is being reset)
  float x_lo, y_lo, cell_x, cell_y, inv_cell_x, inv_cell_y;
  long long* cycles;     // PP_RRT_TIMING: SM cycles of CTA 0 per phase (nearest, edge, barrier 1, append, barrier 2, regroup), else null
};

// Arrival = one release-add, waiting = acquire-loads: the CTA's writes (ordered before thread 0 by the block barrier)
// become visible with the add, and what the other CTAs wrote before theirs is visible after the load that sees the
// target -- the same guarantees as fence + atomic + spin + fence, without the two full memory barriers.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned* generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    *generation += 1;
    const unsigned target = *generation * gridDim.x;
#ifdef PP_RRT_FENCE_BARRIER
    __threadfence();                                   // this CTA's writes are visible before it arrives
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {
      if (++spins > (1u << 28)) __trap();              // a CTA that never arrives is a bug, not something to wait out
    }
    __threadfence();                                   // (gpu-scope fence: also drops this SM's stale L1 lines)
#else
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen = 0, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (++spins > (1u << 28)) __trap();              // a CTA that never arrives is a bug, not something to wait out
    } while (seen < target);
#endif
  }
  __syncthreads();
}

__device__ __forceinline__ int grid_cell_of(const RrtPersist& G, float x, float y) {
  const float fx = (x - G.x_lo) * G.inv_cell_x, fy = (y - G.y_lo) * G.inv_cell_y;
  int ix = fx > 0.f ? (fx >= static_cast<float>(kGridSide) ? kGridSide - 1 : static_cast<int>(fx)) : 0;   // (NaN -> 0)
  int iy = fy > 0.f ? (fy >= static_cast<float>(kGridSide) ? kGridSide - 1 : static_cast<int>(fy)) : 0;
  return ix * kGridSide + iy;
}
// conservative lower bound of dist2 between a sample and any node binned into cell (ix, iy), split per axis: d2 of
// the sample to column ix / row iy of the grid; the box is widened by a slack that covers the float32 binning
// arithmetic, border cells extend to infinity (they hold the clamped far nodes).
// bound(cell) = (axis_x[ix] + axis_y[iy]) * 0.9999f
__device__ __forceinline__ float grid_axis_d2(float lo0, float cell, int i, float q) {
  const float sl = 1e-3f * cell + 1e-3f;
  const float lo = lo0 + i * cell - sl, hi = lo0 + (i + 1) * cell + sl;
  float d = 0.f;
  if (i > 0 && q < lo) d = lo - q;
  if (i < kGridSide - 1 && q > hi) d = q - hi;
  return d * d;
}
__device__ __forceinline__ unsigned long long pack_best(float d, int idx) {
  return (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | static_cast<unsigned int>(idx);
}
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, unsigned long long* s_red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o < v ? o : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long r = s_red[0];
#pragma unroll
  for (int w = 1; w < kT / 32; ++w) r = s_red[w] < r ? s_red[w] : r;
  return r;
}

// minimum of (dist2, id) over the sorted records [a, b), this thread taking first, first + stride, ...: kScanDepth
// independent 16-byte loads in flight (measured: 4 / 8 / 16 -> 20.5 / 21.1 / 22.3 ms for config 4: deeper unrolling spills; a dense bucket at the frontier holds
// thousands of nodes and sets the duration of the whole round)
#ifndef PP_RRT_SCAN_DEPTH
#define PP_RRT_SCAN_DEPTH 4
#endif
constexpr int kScanDepth = PP_RRT_SCAN_DEPTH;
__device__ __forceinline__ unsigned long long scan_sorted(const RrtPersist& G, int a, int b, int first, int stride, float sx,
                                                          float sy, unsigned long long best) {
  for (int k0 = a + first; k0 < b; k0 += kScanDepth * stride) {
    float4 nd[kScanDepth];
#pragma unroll
    for (int u = 0; u < kScanDepth; ++u) { const int k = k0 + u * stride; nd[u] = k < b ? __ldcg(&G.sorted[k]) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int u = 0; u < kScanDepth; ++u) {
      const float dx = __fsub_rn(sx, nd[u].x), dy = __fsub_rn(sy, nd[u].y);
      const unsigned long long c = k0 + u * stride < b ? pack_best(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)), __float_as_int(nd[u].z)) : kBestInit;
      best = c < best ? c : best;
    }
  }
  return best;
}

__global__ void __launch_bounds__(kT) rrt_persistent(RrtArgs A, RrtPersist G) {
  __shared__ __align__(16) uint32_t s_win[kT / 32][kWinWords];
  __shared__ unsigned long long s_red[kT / 32];
  extern __shared__ __align__(16) unsigned char rrt_dyn_smem[];
  int* s_start = reinterpret_cast<int*>(rrt_dyn_smem);                                  // [cells + 1] copy of G.cell_start
  unsigned short* s_cells = reinterpret_cast<unsigned short*>(s_start + kGridCells + 4); // non-empty cells, ascending
  __shared__ int s_ncells;
  __shared__ float s_ax[kGridSide], s_ay[kGridSide];       // per-axis box distances of the current sample
  __shared__ unsigned short s_work[kWorkCap];
  __shared__ int s_nwork, s_hit, s_scan[kT / 32], s_next[2], s_mine[kT];
  __shared__ unsigned s_gen;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const pp_query qq = A.queries[0];
  if (tid == 0) s_gen = 0;
  // replicated round state (identical in every CTA)
  int n_nodes = 1, n_samples = 0, n_rounds = 0, n_coll = 0, n_sorted = 0, goal_id = -1;
  if (blockIdx.x == 0 && tid == 0) {                     // node 0 (written by rrt_init) joins its bucket
    atomicAdd(&G.cell_count[grid_cell_of(G, static_cast<float>(qq.start_x), static_cast<float>(qq.start_y))], 1);
    G.tickets[0] = gridDim.x; G.tickets[1] = gridDim.x;  // every CTA's first sample of a round is its own index
  }
  grid_barrier(G.barrier, &s_gen);
  const int max_samples = A.P.rrt_max_samples, max_nodes = A.P.max_nodes;

  long long cyc[6] = {0, 0, 0, 0, 0, 0};
  long long t_mark = clock64();
#define RRT_TICK(slot) do { if (G.cycles && blockIdx.x == 0 && tid == 0) { const long long now_ = clock64(); cyc[slot] += now_ - t_mark; t_mark = now_; } } while (0)
  while (goal_id < 0 && n_samples < max_samples && n_nodes < max_nodes) {
    const int k_round = min(A.K, max_samples - n_samples);
    // ---- phase A ---------------------------------------------------------------------------------------
    // Samples are handed out by ticket (the searches differ in length: with a fixed share per CTA a third of the
    // round was spent waiting at the barrier below).  The next ticket is requested before the work on the current
    // sample starts, so the atomic's round trip is never waited for.
    if (blockIdx.x == 0 && tid == 0) G.tickets[(n_rounds + 1) & 1] = gridDim.x;   // next round's counter: idle now
    unsigned* ticket = &G.tickets[n_rounds & 1];
    int it = 0;
    for (int t = blockIdx.x; t < k_round; ++it) {
      if (tid == 0) s_next[it & 1] = static_cast<int>(atomicAdd(ticket, 1u));
      const long long t_round = G.cycles ? clock64() : 0;
      float sx, sy;
      draw_sample(A, qq, static_cast<long long>(n_samples) + t, &sx, &sy);
      unsigned long long best = kBestInit;
      // the tail (or the whole tree while it is small): eight independent loads in flight per thread
      for (int k0 = n_sorted + tid; k0 < n_nodes; k0 += 8 * kT) {
        float2 nd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int k = k0 + u * kT; nd[u] = k < n_nodes ? __ldcg(&A.nxy[k]) : make_float2(0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k0 + u * kT;
          const float dx = __fsub_rn(sx, nd[u].x), dy = __fsub_rn(sy, nd[u].y);
          const unsigned long long c = k < n_nodes ? pack_best(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)), k) : kBestInit;
          best = c < best ? c : best;
        }
      }
      int c0 = -1;
      const int n_cells = n_sorted > 0 ? s_ncells : 0;
      if (n_cells > 0) {
        if (tid < kGridSide) s_ax[tid] = grid_axis_d2(G.x_lo, G.cell_x, tid, sx);
        else if (tid < 2 * kGridSide) s_ay[tid - kGridSide] = grid_axis_d2(G.y_lo, G.cell_y, tid - kGridSide, sy);
        __syncthreads();
        unsigned long long mine = kBestInit;                           // (lower bound, cell) of the closest non-empty cell
        for (int k = tid; k < n_cells; k += kT) {
          const int c = s_cells[k];
          const unsigned long long v = pack_best((s_ax[c / kGridSide] + s_ay[c % kGridSide]) * 0.9999f, c);
          mine = v < mine ? v : mine;
        }
        c0 = static_cast<int>(block_min_u64(mine, s_red) & 0xffffffffull);
        if (G.cycles && tid == 0) {
          atomicAdd(reinterpret_cast<unsigned long long*>(&G.cycles[13]), static_cast<unsigned long long>(s_start[c0 + 1] - s_start[c0]));
          atomicMax(reinterpret_cast<unsigned long long*>(&G.cycles[14]), static_cast<unsigned long long>(s_start[c0 + 1] - s_start[c0]));
          atomicMax(reinterpret_cast<unsigned long long*>(&G.cycles[15]), static_cast<unsigned long long>(n_nodes - n_sorted));
        }
        best = scan_sorted(G, s_start[c0], s_start[c0 + 1], tid, kT, sx, sy, best);
      }
      best = block_min_u64(best, s_red);
      if (n_cells > 0) {
        const float bound = __uint_as_float(static_cast<unsigned>(best >> 32));
        if (tid == 0) s_nwork = 0;
        __syncthreads();
        for (int k = tid; k < n_cells; k += kT) {
          const int c = s_cells[k];
          if (c != c0 && (s_ax[c / kGridSide] + s_ay[c % kGridSide]) * 0.9999f <= bound) {
            const int slot = atomicAdd(&s_nwork, 1);
            if (slot < kWorkCap) s_work[slot] = static_cast<unsigned short>(c);
          }
        }
        __syncthreads();
        const int nw = s_nwork;
        if (G.cycles && tid == 0) {               // PP_RRT_TIMING: candidate cells / nodes of step (2), all CTAs
          long long m = 0;
          for (int w = 0; w < min(nw, kWorkCap); ++w) m += s_start[s_work[w] + 1] - s_start[s_work[w]];
          atomicAdd(reinterpret_cast<unsigned long long*>(&G.cycles[8]), static_cast<unsigned long long>(nw));
          atomicAdd(reinterpret_cast<unsigned long long*>(&G.cycles[9]), static_cast<unsigned long long>(m));
          atomicMax(reinterpret_cast<unsigned long long*>(&G.cycles[10]), static_cast<unsigned long long>(nw));
          atomicMax(reinterpret_cast<unsigned long long*>(&G.cycles[11]), static_cast<unsigned long long>(m));
          atomicAdd(reinterpret_cast<unsigned long long*>(&G.cycles[12]), 1ull);
        }
        if (nw > 0) {
          if (nw <= kWorkCap) {
#if PP_RRT_WARP_CELLS
            for (int w = warp; w < nw; w += kT / 32) {                 // one warp per candidate cell
              const int c = s_work[w];
              best = scan_sorted(G, s_start[c], s_start[c + 1], lane, 32, sx, sy, best);
            }
#else
            for (int w = 0; w < nw; ++w) {
              const int c = s_work[w];
              best = scan_sorted(G, s_start[c], s_start[c + 1], tid, kT, sx, sy, best);
            }
#endif
          } else {                                                     // (more candidate cells than the list holds: all sorted nodes)
            best = scan_sorted(G, 0, n_sorted, tid, kT, sx, sy, best);
          }
          best = block_min_u64(best, s_red);
        }
      }
      const int near = static_cast<int>(best & 0xffffffffull);
      RRT_TICK(0);
      long long t_near_end = 0;
      if (G.cycles && tid == 0) {
        t_near_end = clock64();
        if (n_rounds < 4096) atomicMax(reinterpret_cast<unsigned long long*>(&G.cycles[16 + 2 * n_rounds]), static_cast<unsigned long long>(t_near_end - t_round));
      }
      // steer (every thread, identical arithmetic) + the edge's poses shared by the warps
      const Steer S = steer_from(A, qq, near, sx, sy);
      if (tid == 0) s_hit = 0;
      __syncthreads();
      if (S.moves && S.n > 0) {
        const int parts = kT / 32;
        const int k_lo = static_cast<int>(static_cast<long long>(S.n) * warp / parts);
        const int k_hi = static_cast<int>(static_cast<long long>(S.n) * (warp + 1) / parts);
        if (k_hi > k_lo && edge_part_hit(A, S, k_lo, k_hi, s_win[warp], lane) && lane == 0) atomicOr(&s_hit, 1);
      }
      __syncthreads();
      if (tid == 0) {
        A.cand_valid[t] = S.moves ? (s_hit ? 2 : 1) : 0;
        A.cand_near[t] = near;
        double* c = A.cand + static_cast<size_t>(t) * 5;
        c[0] = S.nx; c[1] = S.ny; c[2] = S.yaw; c[3] = S.v_next; c[4] = S.step;
      }
      __syncthreads();
      RRT_TICK(1);
      if (G.cycles && tid == 0 && n_rounds < 4096)
        atomicMax(reinterpret_cast<unsigned long long*>(&G.cycles[17 + 2 * n_rounds]), static_cast<unsigned long long>(clock64() - t_near_end));
      t = s_next[it & 1];                                  // (written before this iteration's barriers, rewritten two iterations on)
    }
    grid_barrier(G.barrier, &s_gen);
    RRT_TICK(2);
    // ---- phase B: order-preserving append ---------------------------------------------------------------
    // every CTA ranks all candidates of the round (one ballot scan per kT of them); the nodes a CTA appends are
    // those whose sample index is congruent to its own, one thread per node
    int appended = 0, coll = 0;
    for (int base = 0; base < k_round; base += kT) {
      const int u = base + tid;
      const int v = u < k_round ? __ldcg(&A.cand_valid[u]) : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, v == 1);
      if (lane == 0) s_scan[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int ww = 0; ww < kT / 32; ++ww) {
        const int c = s_scan[ww];
        if (ww < warp) before += c;
        total += c;
      }
      if (u < k_round && u % static_cast<int>(gridDim.x) == static_cast<int>(blockIdx.x))
        s_mine[u / gridDim.x] = v == 1 ? appended + before + __popc(bal & ((1u << lane) - 1u)) : -1;
      appended += total;
      coll += __syncthreads_count(v == 2);                 // (also: s_scan is free again, s_mine is written)
    }
    const int n_mine = blockIdx.x < static_cast<unsigned>(k_round) ? (k_round - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    if (tid < n_mine && s_mine[tid] >= 0) {
      const int t = blockIdx.x + tid * gridDim.x;
      const int id = n_nodes + s_mine[tid];
      const double* c = A.cand + static_cast<size_t>(t) * 5;
      const double x = __ldcg(&c[0]), y = __ldcg(&c[1]);
      A.nx[id] = x; A.ny[id] = y; A.nhd[id] = __ldcg(&c[2]); A.nvel[id] = __ldcg(&c[3]); A.nstep[id] = __ldcg(&c[4]);
      A.nparent[id] = __ldcg(&A.cand_near[t]);
      const float2 xy = make_float2(static_cast<float>(x), static_cast<float>(y));
      A.nxy[id] = xy;
      atomicAdd(&G.cell_count[grid_cell_of(G, xy.x, xy.y)], 1);
      if (fabs(dm::sub(x, qq.goal_x)) <= A.P.rrt_goal_tolerance && fabs(dm::sub(y, qq.goal_y)) <= A.P.rrt_goal_tolerance)
        atomicMin(G.goal_id, id);         // the first appended node (lowest sample index) in the goal box
    }
    RRT_TICK(3);
    grid_barrier(G.barrier, &s_gen);
    RRT_TICK(4);
    n_nodes += appended;
    n_samples += k_round;
    n_rounds += 1;
    n_coll += coll;
    {
      const int gid = __ldcg(G.goal_id);
      goal_id = gid == 0x7fffffff ? -1 : gid;
    }
    // ---- regroup the nodes by bucket when the tail has grown ------------------------------------------------
    if (goal_id < 0 && n_nodes >= kGridFrom && n_nodes - n_sorted > kTailMax) {
      if (blockIdx.x == 0) {                       // exclusive scan of the per-cell counts (16 cells per thread)
        constexpr int kPer = kGridCells / kT;
        int sum = 0;
#pragma unroll 8
        for (int i = 0; i < kPer; ++i) sum += __ldcg(&G.cell_count[tid * kPer + i]);
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += o;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int run = incl - sum;
        for (int w = 0; w < warp; ++w) run += s_scan[w];
        for (int i = 0; i < kPer; ++i) {
          G.cell_start[tid * kPer + i] = run;
          G.cell_cursor[tid * kPer + i] = 0;
          run += __ldcg(&G.cell_count[tid * kPer + i]);
        }
        if (tid == kT - 1) G.cell_start[kGridCells] = run;
      }
      grid_barrier(G.barrier, &s_gen);
      for (int id = blockIdx.x * kT + tid; id < n_nodes; id += gridDim.x * kT) {
        const float2 xy = __ldcg(&A.nxy[id]);
        const int c = grid_cell_of(G, xy.x, xy.y);
        const int pos = __ldcg(&G.cell_start[c]) + atomicAdd(&G.cell_cursor[c], 1);
        G.sorted[pos] = make_float4(xy.x, xy.y, __int_as_float(id), 0.f);
      }
      grid_barrier(G.barrier, &s_gen);
      n_sorted = n_nodes;
      // every CTA refreshes its copy of the offsets and compacts the non-empty cells (ascending) behind it
#pragma unroll 8
      for (int c = tid; c <= kGridCells; c += kT) s_start[c] = __ldcg(&G.cell_start[c]);
      if (tid == 0) s_ncells = 0;
      __syncthreads();
      {
        constexpr int kPer = kGridCells / kT;                  // consecutive cells per thread
        int cnt = 0;
        for (int i = 0; i < kPer; ++i) cnt += s_start[tid * kPer + i + 1] > s_start[tid * kPer + i] ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += o;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int pos = incl - cnt;
        for (int w = 0; w < warp; ++w) pos += s_scan[w];
        for (int i = 0; i < kPer; ++i)
          if (s_start[tid * kPer + i + 1] > s_start[tid * kPer + i]) s_cells[pos++] = static_cast<unsigned short>(tid * kPer + i);
        if (tid == kT - 1) s_ncells = pos;
      }
      __syncthreads();
      RRT_TICK(5);
    }
  }
  if (G.cycles && blockIdx.x == 0 && tid == 0)
    for (int c = 0; c < 6; ++c) G.cycles[c] = cyc[c];
  if (blockIdx.x == 0 && tid == 0) {
    RrtState st;
    st.n_nodes = n_nodes; st.n_samples = n_samples; st.n_rounds = n_rounds; st.n_coll = n_coll; st.goal_id = goal_id; st.done = 1;
    A.state[0] = st;
  }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(kT) rrt_finalize(RrtArgs A) {
  __shared__ unsigned long long s_dig[kT / 32];
  __shared__ int s_len;
  const int q = blockIdx.x;
  if (q >= A.n_queries) return;
  const RrtState st = A.state[q];
  const size_t nb = static_cast<size_t>(q) * A.NC;
  unsigned long long dig = 0;
  for (int id = threadIdx.x; id < st.n_nodes; id += kT) {
    const int par = A.nparent[nb + id];
    const double x = A.nx[nb + id], y = A.ny[nb + id];
    const double k[6] = {x, y, par >= 0 ? A.nx[nb + par] : x, par >= 0 ? A.ny[nb + par] : y, A.nhd[nb + id],
                         A.nvel[nb + id]};
    unsigned long long h = static_cast<unsigned long long>(id) + 1ull;
    for (int c = 0; c < 6; ++c) h = mix64(h ^ static_cast<unsigned long long>(__double_as_longlong(k[c])));
    dig += h;
  }
  for (int off = 16; off > 0; off >>= 1) dig += __shfl_xor_sync(0xffffffffu, dig, off);
  if ((threadIdx.x & 31) == 0) s_dig[threadIdx.x >> 5] = dig;
  __syncthreads();
  int* ids = A.out_path_ids + static_cast<size_t>(q) * A.cap_path;
  if (threadIdx.x == 0) {
    int len = 0;
    for (int id = st.goal_id; id >= 0; id = A.nparent[nb + id]) len++;
    s_len = len;
    // the chain is written back to front so that the output runs start -> goal
    int k = len - 1;
    for (int id = st.goal_id; id >= 0; id = A.nparent[nb + id], --k)
      if (k < A.cap_path) ids[k] = id;
  }
  __syncthreads();
  const int len = s_len;
  double* path = A.out_path + static_cast<size_t>(q) * A.cap_path * 2;
  double* prof = A.out_profile + static_cast<size_t>(q) * A.cap_path * 3;
  for (int k = threadIdx.x; k < len && k < A.cap_path; k += kT) {
    const int id = ids[k];
    path[2 * k] = A.nx[nb + id];
    path[2 * k + 1] = A.ny[nb + id];
    if (k + 1 < len && k + 1 < A.cap_path) {   // profile in the form of LP:576-580
      const int nid = ids[k + 1];
      prof[k] = dm::div(dm::sub(A.nhd[nb + nid], A.nhd[nb + id]), dm::div(A.P.time_step_ms, 1000.0));
      prof[A.cap_path + k] = A.P.time_step_ms;
      prof[2 * A.cap_path + k] = A.nvel[nb + nid];
    }
  }
  if (threadIdx.x == 0) {
    unsigned long long d = 0;
    for (int w = 0; w < kT / 32; ++w) d += s_dig[w];
    RrtHdr h;
    h.status = st.goal_id >= 0 ? PP_PLAN_FOUND : PP_PLAN_CAP;
    h.n_rounds = st.n_rounds;
    h.n_nodes = st.n_nodes;
    h.n_samples = st.n_samples;
    h.n_coll = st.n_coll;
    h.n_path = len;
    h.digest = d;
    A.hdr[q] = h;
  }
}

}  // namespace

}  // namespace pp

struct RrtWorkspace {
  void* grid = nullptr;   size_t grid_bytes = 0;      // bucket grid of the persistent single-query kernel
  void* nodes = nullptr;  size_t nodes_bytes = 0;
  void* round = nullptr;  size_t round_bytes = 0;
  void* out = nullptr;    size_t out_bytes = 0;
  pp_query* d_queries = nullptr; int queries_cap = 0;
  int* d_active = nullptr;
  int* h_active = nullptr;
  std::vector<unsigned char> h_out;
};

namespace pp {

void rrt_free(pp_handle* h) {
  if (!h->rrt_ws) return;
  cudaFree(h->rrt_ws->grid);
  cudaFree(h->rrt_ws->nodes);
  cudaFree(h->rrt_ws->round);
  cudaFree(h->rrt_ws->out);
  cudaFree(h->rrt_ws->d_queries);
  cudaFree(h->rrt_ws->d_active);
  cudaFreeHost(h->rrt_ws->h_active);
  delete h->rrt_ws;
  h->rrt_ws = nullptr;
}

static int grow_buf(void** p, size_t* have, size_t want) {
  if (want <= *have) return PP_OK;
  cudaFree(*p);
  *p = nullptr; *have = 0;
  PP_CUDA(cudaMalloc(p, want));
  *have = want;
  return PP_OK;
}
static size_t al(size_t b) { return (b + 255) & ~size_t(255); }

int rrt_plan_batch(pp_handle* h, int n, const pp_query* qs, pp_result* rs) {
  const pp_params& P = h->P;
  PP_REQUIRE(P.rrt_samples_per_round >= 1 && P.rrt_samples_per_round <= kMaxK,
             "pp_rrt_plan: rrt_samples_per_round must be in [1, 1024]");
  PP_REQUIRE(P.max_nodes >= 1 && P.rrt_max_samples >= 0, "pp_rrt_plan: bad caps");
  for (int k = 0; k < n; ++k)
    PP_REQUIRE(rs[k].cap_nodes >= 0 && rs[k].cap_path >= 0 && rs[k].cap_expansions >= 0,
               "pp_rrt_plan: negative capacity in a pp_result");
  if (!h->rrt_ws) {
    h->rrt_ws = new RrtWorkspace();
    PP_CUDA(cudaMalloc(&h->rrt_ws->d_active, sizeof(int)));
    PP_CUDA(cudaMallocHost(&h->rrt_ws->h_active, sizeof(int)));
  }
  RrtWorkspace* W = h->rrt_ws;
  cudaStream_t s = h->stream;
  const int K = P.rrt_samples_per_round;
  const long long nc_ll = static_cast<long long>(std::min<long long>(P.max_nodes, static_cast<long long>(P.rrt_max_samples) + 1)) + K + 1;
  const int NC = static_cast<int>(nc_ll);
  int cap_path = 2;
  for (int k = 0; k < n; ++k) cap_path = std::max(cap_path, rs[k].cap_path);
  cap_path = std::min(cap_path, 1 << 20);

  // node storage
  const size_t per_q = static_cast<size_t>(NC);
  const size_t nodes_total = al(sizeof(double) * per_q * n) * 5 + al(sizeof(int) * per_q * n) + al(sizeof(float2) * per_q * n);
  int rc = grow_buf(&W->nodes, &W->nodes_bytes, nodes_total);
  if (rc) return rc;
  const size_t round_total = al(sizeof(unsigned long long) * K * static_cast<size_t>(n)) + al(sizeof(double) * 5 * K * static_cast<size_t>(n)) +
                             2 * al(sizeof(int) * K * static_cast<size_t>(n)) + al(sizeof(RrtState) * n);
  rc = grow_buf(&W->round, &W->round_bytes, round_total);
  if (rc) return rc;
  const size_t hdr_b = al(sizeof(RrtHdr) * n), path_b = al(sizeof(double) * 2 * cap_path * static_cast<size_t>(n)),
               pid_b = al(sizeof(int) * cap_path * static_cast<size_t>(n)), prof_b = al(sizeof(double) * 3 * cap_path * static_cast<size_t>(n));
  const size_t out_total = hdr_b + path_b + pid_b + prof_b;
  rc = grow_buf(&W->out, &W->out_bytes, out_total);
  if (rc) return rc;
  if (n > W->queries_cap) {
    cudaFree(W->d_queries);
    W->d_queries = nullptr; W->queries_cap = 0;
    PP_CUDA(cudaMalloc(&W->d_queries, sizeof(pp_query) * n));
    W->queries_cap = n;
  }
  PP_CUDA(cudaMemcpyAsync(W->d_queries, qs, sizeof(pp_query) * n, cudaMemcpyHostToDevice, s));
  PP_CUDA(cudaMemsetAsync(W->d_active, 0, sizeof(int), s));

  RrtArgs A;
  A.g = h->grid;
  A.cp = collide_params(h);
  A.P = P;
  A.queries = W->d_queries;
  A.n_queries = n;
  A.K = K;
  A.NC = NC;
  A.cap_path = cap_path;
  // sampling box = the grid's world extent (same float expressions as the oracle)
  const float x_lo = static_cast<float>((P.costmap_height / 2 - P.costmap_height) * P.costmap_res);
  const float x_hi = static_cast<float>((P.costmap_height / 2) * P.costmap_res);
  const float y_lo = static_cast<float>((P.costmap_width / 2 - P.costmap_width) * P.costmap_res);
  const float y_hi = static_cast<float>((P.costmap_width / 2) * P.costmap_res);
  A.x_lo = x_lo; A.x_rng = x_hi - x_lo; A.y_lo = y_lo; A.y_rng = y_hi - y_lo;
  A.bias = static_cast<float>(P.rrt_goal_bias);
  unsigned char* nbp = static_cast<unsigned char*>(W->nodes);
  const size_t dsz = al(sizeof(double) * per_q * n);
  A.nx = reinterpret_cast<double*>(nbp);
  A.ny = reinterpret_cast<double*>(nbp + dsz);
  A.nhd = reinterpret_cast<double*>(nbp + 2 * dsz);
  A.nvel = reinterpret_cast<double*>(nbp + 3 * dsz);
  A.nstep = reinterpret_cast<double*>(nbp + 4 * dsz);
  A.nparent = reinterpret_cast<int*>(nbp + 5 * dsz);
  A.nxy = reinterpret_cast<float2*>(nbp + 5 * dsz + al(sizeof(int) * per_q * n));
  unsigned char* rb = static_cast<unsigned char*>(W->round);
  A.best = reinterpret_cast<unsigned long long*>(rb);
  rb += al(sizeof(unsigned long long) * K * static_cast<size_t>(n));
  A.cand = reinterpret_cast<double*>(rb);
  rb += al(sizeof(double) * 5 * K * static_cast<size_t>(n));
  A.cand_near = reinterpret_cast<int*>(rb);
  rb += al(sizeof(int) * K * static_cast<size_t>(n));
  A.cand_valid = reinterpret_cast<int*>(rb);
  rb += al(sizeof(int) * K * static_cast<size_t>(n));
  A.state = reinterpret_cast<RrtState*>(rb);
  unsigned char* ob = static_cast<unsigned char*>(W->out);
  A.hdr = reinterpret_cast<RrtHdr*>(ob);
  A.out_path = reinterpret_cast<double*>(ob + hdr_b);
  A.out_path_ids = reinterpret_cast<int*>(ob + hdr_b + path_b);
  A.out_profile = reinterpret_cast<double*>(ob + hdr_b + path_b + pid_b);
  A.n_active = W->d_active;

  PP_CUDA(cudaEventRecord(h->ev[8], s));
  rrt_init<<<n, kT, 0, s>>>(A);
  h->launches++;
  // a lone query runs its whole round loop inside ONE cooperative kernel (grid barriers instead of launches, bucket-grid
  // nearest neighbour); batches keep the three kernels per round below, which handle all queries of a round at once
  static const bool no_persistent = getenv("PP_RRT_PER_ROUND_KERNELS") != nullptr;
  bool persistent = n == 1 && !no_persistent;
  if (persistent) {
    int coop = 0, per_sm = 0;
    PP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    const size_t dyn_smem = sizeof(int) * (kGridCells + 4) + sizeof(unsigned short) * kGridCells;
    PP_CUDA(cudaFuncSetAttribute(rrt_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn_smem)));
    PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rrt_persistent, kT, dyn_smem));
    if (!coop || per_sm < 1) persistent = false;
    else {
      const size_t g_total = al(sizeof(int) * kGridCells) * 2 + al(sizeof(int) * (kGridCells + 1)) + al(sizeof(float4) * per_q) + 512 + (16 + 2 * 4096) * sizeof(long long);
      rc = grow_buf(&W->grid, &W->grid_bytes, g_total);
      if (rc) return rc;
      unsigned char* gp = static_cast<unsigned char*>(W->grid);
      RrtPersist G;
      G.cell_count = reinterpret_cast<int*>(gp);            gp += al(sizeof(int) * kGridCells);
      G.cell_cursor = reinterpret_cast<int*>(gp);           gp += al(sizeof(int) * kGridCells);
      G.cell_start = reinterpret_cast<int*>(gp);            gp += al(sizeof(int) * (kGridCells + 1));
      G.sorted = reinterpret_cast<float4*>(gp);             gp += al(sizeof(float4) * per_q);
      G.goal_id = reinterpret_cast<int*>(gp);
      G.barrier = reinterpret_cast<unsigned*>(gp + 64);
      G.tickets = reinterpret_cast<unsigned*>(gp + 96);
      static const bool timing = getenv("PP_RRT_TIMING") != nullptr;
      G.cycles = timing ? reinterpret_cast<long long*>(gp + 128) : nullptr;
      if (timing) PP_CUDA(cudaMemsetAsync(gp + 128, 0, (16 + 2 * 4096) * sizeof(long long), s));
      G.x_lo = x_lo; G.y_lo = y_lo;
      G.cell_x = (x_hi - x_lo) / kGridSide; G.cell_y = (y_hi - y_lo) / kGridSide;
      G.inv_cell_x = 1.0f / G.cell_x; G.inv_cell_y = 1.0f / G.cell_y;
      PP_CUDA(cudaMemsetAsync(G.cell_count, 0, sizeof(int) * kGridCells, s));
      PP_CUDA(cudaMemsetAsync(G.cell_start, 0, sizeof(int) * (kGridCells + 1), s));
      const int no_goal = 0x7fffffff;                        // (pageable source: staged before the call returns)
      PP_CUDA(cudaMemcpyAsync(G.goal_id, &no_goal, sizeof(int), cudaMemcpyHostToDevice, s));
      PP_CUDA(cudaMemsetAsync(G.barrier, 0, 4, s));
      const int capacity = per_sm * h->sm_count;
      static const int cta_cap = getenv("PP_RRT_CTAS") ? atoi(getenv("PP_RRT_CTAS")) : 0;   // (experiments: fewer CTAs, more tickets each)
      const int grid = std::max(1, std::min(std::min(K, capacity), cta_cap > 0 ? cta_cap : capacity));
      void* kargs[2] = {&A, &G};
      PP_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(rrt_persistent), dim3(grid), dim3(kT), kargs, dyn_smem, s));
      h->launches++;
      if (timing) {
        static long long cyc[16 + 2 * 4096];
        PP_CUDA(cudaMemcpyAsync(cyc, G.cycles, sizeof(cyc), cudaMemcpyDeviceToHost, s));
        PP_CUDA(cudaStreamSynchronize(s));
        fprintf(stderr, "[pp_rrt] step (1): mean nodes in the closest cell %.0f (max %lld), longest tail %lld\n",
                cyc[12] ? double(cyc[13]) / cyc[12] : 0.0, cyc[14], cyc[15]);
        long long mx_near = 0, mx_edge = 0;
        for (int r = 0; r < 4096; ++r) { mx_near += cyc[16 + 2 * r]; mx_edge += cyc[17 + 2 * r]; }
        fprintf(stderr, "[pp_rrt] sum over rounds of the slowest CTA: nearest %lld cycles, steer+edge %lld cycles\n", mx_near, mx_edge);
        for (int r = 0; r < 4096; r += 50)
          if (cyc[16 + 2 * r]) fprintf(stderr, "[pp_rrt]   round %4d: slowest nearest %lld, slowest steer+edge %lld\n", r, cyc[16 + 2 * r], cyc[17 + 2 * r]);
        fprintf(stderr, "[pp_rrt] step (2): %lld samples, mean cells %.1f (max %lld), mean nodes %.0f (max %lld)\n", cyc[12],
                cyc[12] ? double(cyc[8]) / cyc[12] : 0.0, cyc[10], cyc[12] ? double(cyc[9]) / cyc[12] : 0.0, cyc[11]);
        const char* names[6] = {"nearest", "steer+edge", "barrier 1", "append", "barrier 2", "regroup"};
        for (int c = 0; c < 6; ++c) fprintf(stderr, "[pp_rrt] CTA 0 %-10s %12lld cycles\n", names[c], cyc[c]);
      }
    }
  }
  const long long max_rounds = (static_cast<long long>(P.rrt_max_samples) + K - 1) / K;
  // a lone query's round has too few samples to fill the GPU with one warp per edge: split each edge's poses
  const int ext_split = static_cast<long long>(n) * K <= 2048 ? kExtSplitMax : 1;
  const long long warps_total = static_cast<long long>(n) * K * ext_split;
  const int extend_blocks = static_cast<int>((warps_total + (kT / 32) - 1) / (kT / 32));
  long long round = 0;
  long long nodes_bound = 1;   // upper bound of any query's tree size
  const int chunk = 16;        // rounds between host checks of the "all done" counter
  while (!persistent && round < max_rounds) {
    for (int c = 0; c < chunk && round < max_rounds; ++c, ++round) {
      // enough node slices to occupy the GPU when there are few queries
      // slices of at least kSliceGrain nodes: a lone query's early rounds (a few thousand nodes) are spread over
      // many SMs instead of one CTA scanning a whole 1024-node tile for all K samples
      const long long want_ctas = std::max<long long>(1, (4LL * h->sm_count + n - 1) / n);
      long long slices = std::min<long long>(want_ctas, (nodes_bound + kSliceGrain - 1) / kSliceGrain);
      if (slices < 1) slices = 1;
      int slice_nodes = static_cast<int>((nodes_bound + slices - 1) / slices);
      slice_nodes = ((slice_nodes + kSliceGrain - 1) / kSliceGrain) * kSliceGrain;
      slices = (nodes_bound + slice_nodes - 1) / slice_nodes;
      dim3 grid_nn(static_cast<unsigned>(slices), static_cast<unsigned>(n));
      if (K <= kT) rrt_nearest<1><<<grid_nn, kT, 0, s>>>(A, slice_nodes);
      else if (K <= 2 * kT) rrt_nearest<2><<<grid_nn, kT, 0, s>>>(A, slice_nodes);
      else rrt_nearest<4><<<grid_nn, kT, 0, s>>>(A, slice_nodes);
      rrt_extend<<<extend_blocks, kT, 0, s>>>(A, ext_split);
      rrt_append<<<n, kT, 0, s>>>(A);
      h->launches += 3;
      nodes_bound = std::min<long long>(nodes_bound + K, NC);
    }
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(W->h_active, W->d_active, sizeof(int), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    if (*W->h_active <= 0) break;
  }
  rrt_finalize<<<n, kT, 0, s>>>(A);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  PP_CUDA(cudaEventRecord(h->ev[9], s));
  W->h_out.resize(out_total);
  PP_CUDA(cudaMemcpyAsync(W->h_out.data(), W->out, out_total, cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[8], h->ev[9]);
  h->phase_ms[PP_PHASE_PLAN] = ms;

  const unsigned char* hb = W->h_out.data();
  const RrtHdr* hdr = reinterpret_cast<const RrtHdr*>(hb);
  const double* hpath = reinterpret_cast<const double*>(hb + hdr_b);
  const int* hpid = reinterpret_cast<const int*>(hb + hdr_b + path_b);
  const double* hprof = reinterpret_cast<const double*>(hb + hdr_b + path_b + pid_b);
  int ret = PP_OK;
  std::vector<double> tmp;
  std::vector<int> tmpi;
  for (int k = 0; k < n; ++k) {
    pp_result& r = rs[k];
    const RrtHdr& d = hdr[k];
    r.status = d.status;
    r.n_expansions = d.n_rounds;
    r.n_nodes = d.n_nodes;
    r.n_open = d.n_nodes;
    r.n_closed = 0;
    r.n_candidates = d.n_samples;
    r.n_collisions = d.n_coll;
    r.n_path = d.n_path;
    r.n_profile = d.n_path > 1 ? d.n_path - 1 : 0;
    r.tree_digest = d.digest;
    if (d.n_path > r.cap_path && (r.path_xy || r.path_node_ids)) ret = PP_ERR_CAPACITY;
    const int np = std::min(d.n_path, std::min(r.cap_path, cap_path));
    if (r.path_xy) std::memcpy(r.path_xy, hpath + static_cast<size_t>(k) * cap_path * 2, sizeof(double) * 2 * np);
    if (r.path_node_ids) std::memcpy(r.path_node_ids, hpid + static_cast<size_t>(k) * cap_path, sizeof(int) * np);
    const int nm = std::min(r.n_profile, std::min(r.cap_path, cap_path));
    const double* pr = hprof + static_cast<size_t>(k) * cap_path * 3;
    if (r.yaw_rates) std::memcpy(r.yaw_rates, pr, sizeof(double) * nm);
    if (r.durations) std::memcpy(r.durations, pr + cap_path, sizeof(double) * nm);
    if (r.speeds) std::memcpy(r.speeds, pr + 2 * cap_path, sizeof(double) * nm);
    if (r.node_state || r.node_parent || r.node_closed_seq) {
      if (d.n_nodes > r.cap_nodes) ret = PP_ERR_CAPACITY;
      const int nn = std::min(d.n_nodes, r.cap_nodes);
      const size_t off = static_cast<size_t>(k) * NC;
      tmpi.resize(nn);
      PP_CUDA(cudaMemcpyAsync(tmpi.data(), A.nparent + off, sizeof(int) * nn, cudaMemcpyDeviceToHost, s));
      if (r.node_state) {
        tmp.resize(static_cast<size_t>(5) * nn);
        const double* src[5] = {A.nx, A.ny, A.nhd, A.nvel, A.nstep};
        for (int c = 0; c < 5; ++c)
          PP_CUDA(cudaMemcpyAsync(tmp.data() + static_cast<size_t>(c) * nn, src[c] + off, sizeof(double) * nn,
                                  cudaMemcpyDeviceToHost, s));
      }
      PP_CUDA(cudaStreamSynchronize(s));
      for (int id = 0; id < nn; ++id) {
        if (r.node_parent) r.node_parent[id] = tmpi[id];
        if (r.node_closed_seq) r.node_closed_seq[id] = -1;
        if (r.node_state) {
          double* o = r.node_state + 8 * static_cast<size_t>(id);
          const int par = tmpi[id];
          o[0] = tmp[id]; o[1] = tmp[nn + id];
          o[2] = par >= 0 && par < nn ? tmp[par] : o[0];
          o[3] = par >= 0 && par < nn ? tmp[nn + par] : o[1];
          o[4] = tmp[2 * static_cast<size_t>(nn) + id]; o[5] = tmp[3 * static_cast<size_t>(nn) + id];
          o[6] = tmp[4 * static_cast<size_t>(nn) + id]; o[7] = 0.0;
        }
      }
    }
  }
  if (ret == PP_ERR_CAPACITY) set_last_error("pp_rrt_plan: a caller-provided output array is too small");
  return ret;
}

}  // namespace pp
