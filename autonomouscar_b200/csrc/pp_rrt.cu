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
  const size_t nb = static_cast<size_t>(q) * A.NC + near;
  const double px = A.nx[nb], py = A.ny[nb], ph = A.nhd[nb], pv = A.nvel[nb];
  // steer: heading toward the sample, limited to +-heading_diff per step
  const double th = dm::atan2_(dm::sub(static_cast<double>(sy), py), dm::sub(static_cast<double>(sx), px));
  double d = dm::wrap_pi(dm::sub(th, ph));
  if (d > A.P.heading_diff) d = A.P.heading_diff;
  if (d < -A.P.heading_diff) d = -A.P.heading_diff;
  const double yaw = dm::add(ph, d);
  double v_next = dm::add(pv, dm::div(dm::mul(qq.max_accel, A.P.time_step_ms), 1000.0));
  if (v_next > qq.max_speed) v_next = qq.max_speed;
  const double sum_v = dm::add(pv, v_next);                                  // LP:155
  const double step = dm::div(dm::mul(sum_v, A.P.time_step_ms), 1000.0);
  int verdict = 0;
  double nx = 0, ny = 0;
  if (step > 0) {
    double sn, cs;
    dm::sincos(yaw, &sn, &cs);
    nx = dm::add(px, dm::mul(step, cs));                                      // LP:156
    ny = dm::add(py, dm::mul(step, sn));                                      // LP:157
    // edge collision, poses k = 1..max(n,1) with n = int(dist / search_res * 2) (LP:308)
    const double dx = dm::sub(nx, px), dy = dm::sub(ny, py);
    const double dist = __dsqrt_rn(dm::add(dm::mul(dx, dx), dm::mul(dy, dy)));
    int n = 0;
    bool hit = false;
    if (dm::trunc_to_int(dm::mul(dm::div(dist, A.P.search_res), 2.0), &n)) {
      if (n < 1) n = 1;
      // the edge's poses are a mini-batch: lane l sets up pose base + l + 1 (broad phase, one
      // thread each), then the warp runs the narrow phase only for the poses that need it
      const int k_lo = static_cast<int>(static_cast<long long>(n) * part / ext_split);        // poses k_lo+1 .. k_hi
      const int k_hi = static_cast<int>(static_cast<long long>(n) * (part + 1) / ext_split);
      for (int base = k_lo; base < k_hi && !hit; base += 32) {
        const int k = base + lane + 1;
        bool need = false;
        NarrowJob job;
        if (k <= k_hi) {
          const double tt = dm::div(static_cast<double>(k), static_cast<double>(n));
          const double ex = dm::add(px, dm::mul(tt, dx)), ey = dm::add(py, dm::mul(tt, dy));
          if (A.cp.mode == PP_COLLISION_ORIENTED) {
            const PoseFx pf = pose_to_fixed(A.cp.lat, A.g.W, A.g.H, A.g.res, ex, ey, yaw);
            need = make_job(A.g, A.cp.lat, pf, 1, &job);
          } else if (A.cp.mode == PP_COLLISION_REF_AABB) {
            hit = aabb_hit(A.g, ex, ey);
          }
        }
        if (A.cp.mode == PP_COLLISION_ORIENTED) hit = warp_collide_jobs(A.g, A.cp.lat, job, need, smem[wib], lane);
        else hit = __any_sync(0xffffffffu, hit);
      }
    }
    verdict = hit ? 2 : 1;
  }
  if (lane == 0 && verdict) atomicMax(&A.cand_valid[slot], verdict);
  if (lane == 0 && part == 0) {
    A.cand_near[slot] = near;
    double* c = A.cand + slot * 5;
    c[0] = nx; c[1] = ny; c[2] = yaw; c[3] = v_next; c[4] = step;
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
  const long long max_rounds = (static_cast<long long>(P.rrt_max_samples) + K - 1) / K;
  // a lone query's round has too few samples to fill the GPU with one warp per edge: split each edge's poses
  const int ext_split = static_cast<long long>(n) * K <= 2048 ? kExtSplitMax : 1;
  const long long warps_total = static_cast<long long>(n) * K * ext_split;
  const int extend_blocks = static_cast<int>((warps_total + (kT / 32) - 1) / (kT / 32));
  long long round = 0;
  long long nodes_bound = 1;   // upper bound of any query's tree size
  const int chunk = 16;        // rounds between host checks of the "all done" counter
  while (round < max_rounds) {
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
