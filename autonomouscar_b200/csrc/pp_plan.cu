// pp_plan.cu -- the reference planner's search loop, one thread block per query.
//
// Replaces Prius::LocalPlanner::planPath and everything it calls
// (catkin_ws/src/local_planner/src/local_planner.cpp, "LP"):
//   planPath LP:25-82 . initializePlanner LP:84-96 . openNode/closeNode LP:98-113 .
//   calcH LP:115-137 . calcW LP:139-144 . getNeighbors LP:146-169 . checkGoalDist LP:171-189 .
//   calcPossibleVelocities LP:191-216 . calcPossibleYaws LP:218-232 . checkForGoal LP:269-287 .
//   interpolateToGoalPoint LP:289-299 . interpolatePoints LP:301-316 . markVisited LP:318-326 .
//   calcPathMsg LP:503-540 . calcMPMessage LP:542-583 . GraphNode::operator== graph_node.hpp:24-30 .
//   CheaperCost + std::priority_queue local_planner.hpp:103-111
//
// B200 mapping (not a port of the CPU structure):
//   - a query is one CTA for its whole life; the tree lives in HBM/L2 as SoA arrays, the
//     frontier heap in shared memory when it fits (it is the latency-critical structure);
//   - the frontier reproduces libstdc++'s push_heap / pop_heap element for element, so equal
//     costs pop in the reference's order; the sift-up is done by one warp in parallel
//     (each lane owns one ancestor);
//   - the reference's O(N) std::find lookups per candidate (97 % of its CPU time) are
//     answered by a per-query hash set of 64-bit key fingerprints in L2; only a fingerprint
//     hit falls back to the exact brute-force scan, so results are identical to the scan;
//   - candidate generation, collision and cost are one thread per (velocity, yaw) child;
//   - path / profile extraction are block-wide first-match / last-match reductions that
//     reproduce the list scan order (open list in insertion order, then closed list in
//     closing order).
#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pp_common.cuh"

namespace pp {

namespace {

constexpr int kTMax = 256;         // threads per CTA: 256 for a lone query (latency), 128 in batches (more CTAs per SM)
constexpr int kWarpsMax = kTMax / 32;
#define kT (static_cast<int>(blockDim.x))
#define kWarps (static_cast<int>(blockDim.x) >> 5)
constexpr int kMaxVelRaw = 1024;   // raw velocity lattice points per expansion (LP:200-203) the kernel walks
constexpr int kMaxCand = 8192;     // children per expansion the kernel supports
constexpr int kIntMax = 0x7fffffff;
constexpr int kFxMaxPoses = 512;   // longest path the pointer-walk extraction handles (longer: exact scan walk)
constexpr int kFxTable = 1024;     // slots of its pose table (power of two)
constexpr unsigned short kFxShared = 0xFFFFu;

struct PlanHdr {                   // per query, device -> host
  int status, n_expansions, n_nodes, n_open, n_closed, n_path, n_profile, n_candidates, n_collisions;
  int truncated;                   // a caller-visible array was too small
  int reran;                       // 1 if a cost tie forced the exact-heap re-run
  int pad_;
  long long pool_off;              // first slot of this query in the output pools
  unsigned long long digest;
  long long cyc[8];                // SM cycles per phase (thread 0): init, close, eval, commit, push, close2, extract, total
};

struct PlanArgs {
  GridDev g;              // the handle's grid, or grid 0 of its bank
  GridBank bank;          // bank.n > 0: query.reserved names the grid (pp_plan_batch_banked)
  CollideParams cp;
  pp_params P;
  const pp_query* queries;
  int n_queries;
  int NC;                 // slot capacity per workspace
  int hash_size;          // power of two >= 2 * NC: the room the workspace reserves for the fingerprint set
  int hash_init;          // power of two <= hash_size: what a query starts with (cleared per query); the set is
                          // rebuilt four times as large whenever the tree could outgrow half of it
  int max_children;       // children one expansion can add (host bound)
  int heap_smem_entries;  // frontier entries kept in shared memory (>= 1)
  int ws_per_query;       // workspace index = query index (node arrays wanted) or CTA index
  int want_path_ids;
  int cap_path, cap_exp;
  int timing;             // accumulate per-phase SM cycles (PP_PLAN_TIMING=1)
  int fast_frontier;      // 1: try the unordered-array frontier first (exact; falls back on cost ties)
  size_t ws_stride;       // bytes
  unsigned char* ws_base;
  PlanHdr* hdr;
  // path / profile outputs are packed: a query reserves min(n_path, cap_path) slots of the
  // pools with one atomicAdd, so the host copies back only what was produced
  double* out_path;       // pool [pool_cap][2]
  int* out_path_ids;      // pool [pool_cap]
  double* out_profile;    // 3 planes of [pool_cap]: yaw rates, durations, speeds
  unsigned long long* pool_counter;
  long long pool_cap;
  // record mode (pp_plan_batch_dev): query q writes a fixed-stride record at rec_base + q * rec_stride
  // (pp_plan_record_header | path_xy | yaw_rates | durations | speeds | path_node_ids, each padded to cap_path)
  // instead of reserving pool slots; the records stay on the device for the result gather
  unsigned char* rec_base;
  size_t rec_stride;
  int* out_exp;           // [n][cap_exp] or null
  int* work_counter;
  uint8_t* visited;       // A12 debug bitmap (single-query plans only) or null
};

// SoA view of one workspace
struct WS {
  double *cx, *cy, *px, *py, *hd, *vel, *g, *cost;
  unsigned long long* fp;     // key fingerprint per slot
  unsigned long long* hash;   // open-addressing set of fingerprints (0 = empty)
  double* heap_cost_g;
  int* closed_seq;
  int* heap_id_g;
  int* closed_list;           // slot id by closing order (m_tree_closed)
  int* rev_last;              // path extraction scratch
  unsigned long long* cmin;   // fast frontier: bit pattern of the smallest cost in every chunk of 32 entries
};

__host__ __device__ inline size_t ws_bytes(int NC, int hash_size) {
  size_t b = 0;
  b += sizeof(double) * 8 * NC;
  b += sizeof(unsigned long long) * NC;
  b += sizeof(unsigned long long) * hash_size;
  b += sizeof(double) * 2 * NC;       // global heap entries (16 B each); reused to stage the reversed path
  b += sizeof(int) * NC * 4;          // closed_seq, path scratch, closed_list, path scratch
  b += sizeof(unsigned long long) * (NC / 32 + 2);   // chunk minima of the fast frontier
  return (b + 255) & ~size_t(255);
}

__device__ __forceinline__ WS ws_view(unsigned char* base, int NC, int hash_size) {
  WS w;
  double* d = reinterpret_cast<double*>(base);
  w.cx = d; w.cy = d + NC; w.px = d + 2 * NC; w.py = d + 3 * NC;
  w.hd = d + 4 * NC; w.vel = d + 5 * NC; w.g = d + 6 * NC; w.cost = d + 7 * NC;
  w.fp = reinterpret_cast<unsigned long long*>(d + 8 * NC);
  w.hash = w.fp + NC;
  w.heap_cost_g = reinterpret_cast<double*>(w.hash + hash_size);
  w.closed_seq = reinterpret_cast<int*>(w.heap_cost_g + 2 * NC);
  w.heap_id_g = w.closed_seq + NC;
  w.closed_list = w.heap_id_g + NC;
  w.rev_last = w.closed_list + NC;
  w.cmin = reinterpret_cast<unsigned long long*>(w.rev_last + NC);
  return w;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long dbits(double d) {
  return static_cast<unsigned long long>(__double_as_longlong(d));
}
// fingerprint of the 6 key doubles; -0.0 and +0.0 compare equal, so canonicalise first
__device__ __forceinline__ unsigned long long key_fingerprint(const double k[6]) {
  unsigned long long h = 0x5851F42D4C957F2Dull;
#pragma unroll
  for (int i = 0; i < 6; ++i) h = mix64(h ^ dbits(__dadd_rn(k[i], 0.0)));
  return h | 1ull;  // 0 is the empty marker of the hash set
}
// 32-bit hash of a point for the extraction scans (== treats -0.0 and +0.0 as equal)
__device__ __forceinline__ uint32_t point_hash32(double x, double y) {
  return static_cast<uint32_t>(mix64(dbits(__dadd_rn(x, 0.0)) ^ mix64(dbits(__dadd_rn(y, 0.0)))) >> 32);
}
// slot of a point in the pointer-walk extraction's pose table: equal points (== on doubles, so -0.0 and +0.0
// alike) must land in the same slot, nothing more -- a few shifted xors of the raw bits instead of two mix64
__device__ __forceinline__ uint32_t point_slot(double x, double y, uint32_t mask) {
  const unsigned long long a = dbits(__dadd_rn(x, 0.0)), b = dbits(__dadd_rn(y, 0.0));
  const unsigned long long m = (a ^ (a >> 23)) + (b ^ (b >> 17)) * 0x9E3779B1ull;
  return static_cast<uint32_t>(m ^ (m >> 29)) & mask;
}
// tree digest term of pp_result.tree_digest (same as oracle/planner_ref.h node_digest)
__device__ __forceinline__ unsigned long long node_digest(int id, const double k[6]) {
  unsigned long long h = static_cast<unsigned long long>(id) + 1ull;
#pragma unroll
  for (int i = 0; i < 6; ++i) h = mix64(h ^ dbits(k[i]));
  return h;
}

// ---------------------------------------------------------------------------------------
// block-wide reductions (all threads must call)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int block_min(int v, int* s_red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = s_red[0];
  for (int w = 1; w < kWarps; ++w) r = min(r, s_red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ int block_max(int v, int* s_red) { return -block_min(-v, s_red); }
__device__ __forceinline__ int block_sum(int v, int* s_red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
  for (int w = 0; w < kWarps; ++w) r += s_red[w];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------
// the frontier: libstdc++ binary heap, comp(a, b) = a.cost > b.cost (local_planner.hpp:103-109)
// ---------------------------------------------------------------------------------------
struct __align__(16) HeapEnt {
  double cost;
  int id;
  int pad;
};
// The top `m` entries (levels near the root, touched by every operation) live in shared
// memory, deeper levels in the query's workspace in L2.  m = capacity for a single query
// (lowest latency); a batch uses a small m so that several CTAs share an SM and overlap each
// other's serial heap phases.
struct Heap {
  HeapEnt* s;
  HeapEnt* g;
  int m;
};
template <bool kAllSmem>
__device__ __forceinline__ HeapEnt hget(const Heap& h, int pos) {
  if (kAllSmem) return h.s[pos];
  return pos < h.m ? h.s[pos] : h.g[pos];
}
template <bool kAllSmem>
__device__ __forceinline__ void hset(const Heap& h, int pos, const HeapEnt& v) {
  if (kAllSmem || pos < h.m) h.s[pos] = v;
  else h.g[pos] = v;
}

// std::__push_heap(first, hole, top = 0, value): executed by one full warp.  Lane k >= 1 owns
// the k-th ancestor of the hole; all ancestors that compare "less" (cost > value) shift down
// one level at once.  One 16-byte load and at most one 16-byte store per lane.
template <bool kAllSmem>
__device__ __forceinline__ void heap_sift_up_warp(const Heap& h, int hole, double vcost, int vid, int lane) {
  const unsigned j = static_cast<unsigned>(hole) + 1u;  // 1-based position
  bool moves = false;
  HeapEnt mine;
  mine.cost = 0; mine.id = 0; mine.pad = 0;
  if (lane >= 1 && (j >> (lane - 1)) >= 2u) {            // the (lane-1)-th ancestor is not the root
    mine = hget<kAllSmem>(h, static_cast<int>(j >> lane) - 1);
    moves = mine.cost > vcost;
  }
  const unsigned stay = __ballot_sync(0xffffffffu, !moves) & ~1u;  // lanes >= 1 that stop the loop
  const int stop = __ffs(stay) - 1;                               // first k with !moves (always exists)
  if (lane >= 1 && lane < stop) hset<kAllSmem>(h, static_cast<int>(j >> (lane - 1)) - 1, mine);
  if (lane == 0) {
    HeapEnt v;
    v.cost = vcost; v.id = vid; v.pad = 0;
    hset<kAllSmem>(h, static_cast<int>(j >> (stop - 1)) - 1, v);
  }
  __syncwarp();
}

// priority_queue::pop by one warp: std::pop_heap (__pop_heap + __adjust_heap) then pop_back.
// `size` is the warp-uniform current size; returns the new size.
template <bool kAllSmem>
__device__ __forceinline__ int heap_pop_warp(const Heap& h, int size, int lane) {
  if (size <= 1) return 0;
  const int len = size - 1;
  const HeapEnt v = hget<kAllSmem>(h, len);
  int hole = 0;
  if (lane == 0) {
    int child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      const HeapEnt r = hget<kAllSmem>(h, child), l = hget<kAllSmem>(h, child - 1);
      if (r.cost > l.cost) { child--; hset<kAllSmem>(h, hole, l); }
      else hset<kAllSmem>(h, hole, r);
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      hset<kAllSmem>(h, hole, hget<kAllSmem>(h, child - 1));
      hole = child - 1;
    }
  }
  hole = __shfl_sync(0xffffffffu, hole, 0);
  __syncwarp();
  heap_sift_up_warp<kAllSmem>(h, hole, v.cost, v.id, lane);
  return len;
}

// ---------------------------------------------------------------------------------------
// reference arithmetic (deterministic math; operation order as in the cited lines)
// ---------------------------------------------------------------------------------------
struct QueryD {
  double gx, gy, gs, max_speed, max_accel, start_speed;
};

// Values that the reference recomputes for every child but that only depend on the query
// (common sub-expressions: same operations, same bits).
struct QueryConst {
  double accel_step;   // max_accel * time_step_ms / 1000          (LP:173)
  double max_dist;     // sqrt(goal_x^2 + goal_y^2)                (LP:124)
};

// LP:171-189.  (s, c) = sin / cos of `heading`, already computed by the caller for the same angle.
__device__ __forceinline__ bool check_goal_dist(const pp_params& P, const QueryD& Q, const QueryConst& K, double cx,
                                                double cy, double s, double c, double vel) {
  double next_vel = dm::add(vel, K.accel_step);
  if (next_vel > Q.max_speed) next_vel = Q.max_speed;
  const double avg = dm::div(dm::add(next_vel, vel), 2.0);
  const double step = dm::div(dm::mul(avg, P.time_step_ms), 1000.0);
  const double x = dm::add(cx, dm::mul(step, c));
  const double y = dm::add(cy, dm::mul(step, s));
  const double dx = dm::sub(x, Q.gx), dy = dm::sub(y, Q.gy);
  const double dist = __dsqrt_rn(dm::add(dm::mul(dx, dx), dm::mul(dy, dy)));
  const double t = dm::div(dm::sub(vel, Q.gs), Q.max_accel);
  const double d_decel = dm::sub(dm::mul(vel, t), dm::div(dm::mul(Q.max_accel, dm::mul(t, t)), 2.0));
  return dm::add(dist, P.goal_pos_tolerance) <= d_decel;
}

// LP:115-137
__device__ __forceinline__ double calc_h(const pp_params& P, const QueryD& Q, const QueryConst& K, double cx,
                                         double cy, double heading, double s, double c, double vel) {
  double speed = vel;
  if (speed == 0) speed = 0.01;
  const double ex = dm::sub(cx, Q.gx), ey = dm::sub(cy, Q.gy);
  const double dist_to_goal = __dsqrt_rn(dm::add(dm::mul(ex, ex), dm::mul(ey, ey)));
  const double dist_h = dm::mul(dm::div(dist_to_goal, K.max_dist), 100.0);
  const double yaw_to_goal = dm::atan2_(dm::sub(Q.gy, cy), dm::sub(Q.gx, cx));
  const double yaw_h = dm::mul(dm::div(fabs(dm::sub(heading, yaw_to_goal)), 6.283185307179586), 100.0);
  double speed_h = dm::sub(100.0, dm::mul(dm::div(speed, P.max_velocity), 100.0));
  if (check_goal_dist(P, Q, K, cx, cy, s, c, vel)) speed_h = dm::mul(fabs(dm::sub(speed, Q.gs)), 100.0);
  return fabs(dm::add(dm::add(dist_h, yaw_h), speed_h));
}

// one comparison of checkForGoal (LP:276-283): abs(d) <= tol under the double or the int overload of abs
__device__ __forceinline__ bool goal_tol_ok(double d, double tol, bool int_abs) {
  if (!int_abs) return fabs(d) <= tol;
  int t;
  if (!dm::trunc_to_int(d, &t)) return false;
  return static_cast<double>(abs(t)) <= tol;
}

struct Shared {
  int n_nodes, n_closed, heap_size, n_exp, status, dup_seen, cur_id, n_vel, n_yaw, n_cand_total, n_coll_total;
  int goal_hit, path_len, truncated, done;
  int q;
  int hash_cur;    // current size of the fingerprint set (power of two, hash_init .. hash_size)
  int fast;        // frontier mode of this attempt: 1 = unordered array + block argmin, 0 = exact binary heap
  int tie_seen;    // fast mode: some pop saw its minimum cost on more than one entry
  int top_pos, top_id;
  int fx_len, fx_bad;   // pointer-walk extraction: path length, 1 = cannot be used
  int fx_slow;          // diagnostics: this query's path was extracted by the exact scan walk
  long long pool_off;
  double goal_node[6];  // child.x child.y parent.x parent.y heading velocity of the terminal node
  double scratch_key[6];  // key broadcast for the exact-scan fallback
  long long cyc[8];
  GridDev gq;      // banked batches: this query's grid (pointers of A.g moved to grid query.reserved)
};

// scan key of a slot: open slot -> id ; closed slot -> n_nodes + closed_seq (open list by id, then closed
// list by closing order: the order LP:511-526 scans in)
// slot id of a scan key
__device__ __forceinline__ int slot_of_key(const WS& w, int n_nodes, int key) {
  return key < n_nodes ? key : w.closed_list[key - n_nodes];
}

// exact brute-force lookup of LP:56-57 for one key: first OPEN slot equal to the key (lowest
// id) and whether ANY closed slot equals it.  Block-wide.
__device__ __forceinline__ void scan_exact(const WS& w, int n_nodes, const double k[6], int* first_open,
                                           int* any_closed, int* s_red) {
  int fo = kIntMax, ac = 0;
  for (int id = threadIdx.x; id < n_nodes; id += kT) {
    if (w.cx[id] == k[0] && w.cy[id] == k[1] && w.px[id] == k[2] && w.py[id] == k[3] && w.hd[id] == k[4] &&
        w.vel[id] == k[5]) {
      if (w.closed_seq[id] < 0) fo = min(fo, id);
      else ac = 1;
    }
  }
  *first_open = block_min(fo, s_red);
  *any_closed = block_max(ac, s_red);
}

__device__ __forceinline__ bool hash_contains(const WS& w, int hash_size, unsigned long long fp) {
  unsigned pos = static_cast<unsigned>(fp >> 17) & (hash_size - 1);
  for (;;) {
    const unsigned long long e = w.hash[pos];
    if (e == 0ull) return false;
    if (e == fp) return true;
    pos = (pos + 1) & (hash_size - 1);
  }
}
__device__ __forceinline__ void hash_insert(const WS& w, int hash_size, unsigned long long fp) {
  unsigned pos = static_cast<unsigned>(fp >> 17) & (hash_size - 1);
  for (;;) {
    const unsigned long long e = atomicCAS(&w.hash[pos], 0ull, fp);
    if (e == 0ull || e == fp) return;
    pos = (pos + 1) & (hash_size - 1);
  }
}

// ---------------------------------------------------------------------------------------
// Fast frontier.  The order in which a priority queue returns its elements does not depend
// on its internal layout as long as every pop has a UNIQUE minimum.  The fast mode therefore
// keeps the frontier as an unordered array (pushes are plain parallel appends) and pops by a
// block-wide argmin; if any pop ever sees its minimum cost on two entries the query is
// re-run from scratch with the exact libstdc++ heap, so results are always the reference's.
//
// Two levels keep the argmin short: w.cmin[k] holds the smallest cost among entries 32 k .. 32 k + 31 (as the bit
// pattern of the double: costs are non-negative, so the patterns order like the values).  A push is an atomicMin
// on its chunk's word; a removal (the last entry moves into the hole) recomputes the two chunks it touches.  A pop
// reads the chunk minima (n / 32 words, one or two per thread), then one warp reads the 32 entries of the winning
// chunk.  The minimum is tied exactly when two chunk minima equal it or two entries of that chunk do.
// (Before: one pass over all n entries per expansion that also found the runner-up for the next pop -- 24 % of the
// kernel's cycles; with the chunk minima in shared memory and cached in registers the 64-register build spilled
// and lost what the shorter pass had won.)
// ---------------------------------------------------------------------------------------
constexpr unsigned long long kCostInf = 0x7FF0000000000000ull;
__device__ __forceinline__ unsigned long long cost_bits(double c) { return static_cast<unsigned long long>(__double_as_longlong(c)); }

template <bool kAllSmem>
__device__ __forceinline__ void fast_find_top(const Heap& h, const WS& w, Shared* S, unsigned long long* s_min) {
  const int n = S->heap_size, nch = (n + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long best = ~0ull;
  for (int k = threadIdx.x; k < nch; k += kT) best = min(best, w.cmin[k]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
  if (lane == 0) s_min[warp] = best;
  __syncthreads();
  unsigned long long gmin = s_min[0];
#pragma unroll
  for (int ww = 1; ww < kWarps; ++ww) gmin = min(gmin, s_min[ww]);
  int cnt = 0, mine = -1;
  for (int k = threadIdx.x; k < nch; k += kT)
    if (w.cmin[k] == gmin) { cnt++; mine = k; }
  if (cnt > 0) S->top_pos = mine;                       // (the chunk for now; any one of them if several: a tie anyway)
  if (cnt > 1) S->tie_seen = 1;
  const int holders = __syncthreads_count(cnt > 0);     // (also orders the writes above before the reads below)
  if (warp == 0) {
    const int k = S->top_pos, idx = 32 * k + lane;
    HeapEnt e;
    e.cost = 0.0; e.id = -1; e.pad = 0;
    if (idx < n) e = hget<kAllSmem>(h, idx);
    const unsigned bal = __ballot_sync(0xffffffffu, idx < n && cost_bits(e.cost) == gmin);
    __syncwarp();
    if (lane == __ffs(bal) - 1) {
      S->top_pos = idx;
      S->top_id = e.id;
      if (__popc(bal) > 1 || holders > 1) S->tie_seen = 1;
    }
  }
  __syncthreads();
}

// frontier.top(): slot id of the cheapest entry (block-wide; all threads get the same value)
template <bool kAllSmem>
__device__ __forceinline__ int frontier_top(const Heap& h, const WS& w, Shared* S, unsigned long long* s_min) {
  if (S->fast) {
    fast_find_top<kAllSmem>(h, w, S, s_min);
    return S->top_id;
  }
  return h.s[0].id;
}

// openNode (LP:98-102) for one node by thread 0 + frontier push by warp 0.  Called by warp 0 only.
template <bool kAllSmem>
__device__ __forceinline__ void open_single_warp0(const WS& w, const Heap& heap, Shared* S, int hash_size,
                                                 const double k[6], double g, double cost, int parent_slot, int lane) {
  const int id = S->n_nodes;
  const int hsize = S->heap_size;
  const int fast = S->fast;
  __syncwarp();
  if (lane == 0) {
    w.cx[id] = k[0]; w.cy[id] = k[1]; w.px[id] = k[2]; w.py[id] = k[3]; w.hd[id] = k[4]; w.vel[id] = k[5];
    w.g[id] = g; w.cost[id] = cost;
    w.closed_seq[id] = -1;
    w.rev_last[id] = parent_slot;   // slot that was being expanded when this node was opened (extraction hint)
    const unsigned long long fp = key_fingerprint(k);
    w.fp[id] = fp;
    if (hash_contains(w, hash_size, fp)) S->dup_seen = 1;
    hash_insert(w, hash_size, fp);
    if (fast) {
      HeapEnt e;
      e.cost = cost; e.id = id; e.pad = 0;
      hset<kAllSmem>(heap, hsize, e);
      atomicMin(&w.cmin[hsize >> 5], cost_bits(cost));
    }
  }
  __syncwarp();
  if (!fast) heap_sift_up_warp<kAllSmem>(heap, hsize, cost, id, lane);
  if (lane == 0) { S->n_nodes = id + 1; S->heap_size = hsize + 1; }
  __syncwarp();
}

// closeNode (LP:104-113): find the first open slot equal to frontier.top() (`top_id`), move it
// to the closed list, pop the frontier.  Block-wide; returns the closed slot id (or -1).
template <bool kAllSmem>
__device__ __forceinline__ int close_top(const WS& w, const Heap& heap, Shared* S, int top_id, int* s_red) {
  int target = top_id;
  if (S->dup_seen) {
    // equal keys may exist: reproduce std::find over m_tree_open (first equal, lowest id)
    double k[6] = {w.cx[top_id], w.cy[top_id], w.px[top_id], w.py[top_id], w.hd[top_id], w.vel[top_id]};
    int fo, ac;
    scan_exact(w, S->n_nodes, k, &fo, &ac, s_red);
    target = fo == kIntMax ? -1 : fo;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int hsize = S->heap_size;
    const int fast = S->fast;
    __syncwarp();
    if (threadIdx.x == 0 && target >= 0) {
      w.closed_seq[target] = S->n_closed;
      w.closed_list[S->n_closed] = target;
      S->n_closed = S->n_closed + 1;
    }
    if (fast) {
      // remove by moving the last entry into the hole, then recompute the minima of the (up to) two chunks touched
      const int pos = S->top_pos, last = hsize - 1, lane = threadIdx.x;
      __syncwarp();
      if (lane == 0 && pos != last) hset<kAllSmem>(heap, pos, hget<kAllSmem>(heap, last));
      __syncwarp();
      const int ka = pos >> 5, kb = last >> 5;
      for (int pass = 0; pass < 2; ++pass) {
        const int k = pass == 0 ? ka : kb;
        if (pass == 1 && kb == ka) break;
        const int idx = 32 * k + lane;
        unsigned long long v = kCostInf;
        if (idx < last) v = cost_bits(hget<kAllSmem>(heap, idx).cost);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) w.cmin[k] = v;
      }
      if (lane == 0) S->heap_size = last;
    } else {
      const int nsize = heap_pop_warp<kAllSmem>(heap, hsize, threadIdx.x);
      if (threadIdx.x == 0) S->heap_size = nsize;
    }
  }
  __syncthreads();
  return target;
}

// kBatchCtas > 0: the batch build -- 128 threads per CTA and a register budget that lets
// kBatchCtas CTAs share an SM (each query is latency-bound on its own; throughput comes from
// overlapping many of them).  0: the lone-query build, 256 threads, no register cap.
// Two batch builds (measured on B200, profiles/README.md): while all queries fit one wave of 6 CTAs per SM the
// 80-register build is faster (fewer spills, 5-8 %); beyond that the 64-register build with 8 CTAs per SM wins by
// 6-22 %: the queries are latency-bound, so more of them in flight per SM pays, and the last wave is fuller.
#define PP_PLAN_BATCH_CTAS 6
#define PP_PLAN_BATCH_HEAP 2048        // frontier entries a CTA of that build keeps in shared memory (32 KB)
#define PP_PLAN_BATCH_CTAS_WIDE 8
#define PP_PLAN_BATCH_HEAP_WIDE 1280   // 20 KB
template <bool kAllSmem, int kBatchCtas, bool kBanked>
__global__ void __launch_bounds__(kBatchCtas ? 128 : kTMax, kBatchCtas ? kBatchCtas : 1) plan_kernel(PlanArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Shared S;
  __shared__ int s_red[kWarpsMax];
  __shared__ int s_scan[kWarpsMax];
  __shared__ unsigned long long s_dig[kWarpsMax];
  __shared__ unsigned long long s_fmin[kWarpsMax];
  __shared__ int s_fpos[kWarpsMax];
  // pointer-walk extraction: direct-mapped table over the path poses' point hashes.  Entry = pose index + 1,
  // 0 = no pose hashes here, kFxShared = more than one does (those few slots fall back to a scan of the path)
  __shared__ unsigned short s_ptab[kFxTable];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const pp_params& P = A.P;
  const int NC = A.NC;

  for (;;) {
    if (tid == 0) S.q = atomicAdd(A.work_counter, 1);
    __syncthreads();
    const int q = S.q;
    __syncthreads();
    if (q >= A.n_queries) return;

    const WS w = ws_view(A.ws_base + static_cast<size_t>(A.ws_per_query ? q : blockIdx.x) * A.ws_stride, NC,
                         A.hash_size);
    Heap heap;
    heap.s = reinterpret_cast<HeapEnt*>(smem_raw);
    heap.g = reinterpret_cast<HeapEnt*>(w.heap_cost_g);
    heap.m = A.heap_smem_entries;
    const pp_query qq = A.queries[q];
    QueryD Q;
    Q.gx = qq.goal_x; Q.gy = qq.goal_y; Q.gs = qq.goal_speed;
    Q.max_speed = qq.max_speed; Q.max_accel = qq.max_accel; Q.start_speed = qq.start_speed;
    QueryConst KQ;
    KQ.accel_step = dm::div(dm::mul(Q.max_accel, P.time_step_ms), 1000.0);
    KQ.max_dist = __dsqrt_rn(dm::add(dm::mul(Q.gx, Q.gx), dm::mul(Q.gy, Q.gy)));
    int* exp_out = A.out_exp ? A.out_exp + static_cast<size_t>(q) * A.cap_exp : nullptr;

    // attempt 0 runs with the fast frontier; a detected cost tie re-runs the query with the exact heap
    const long long t_begin = A.timing ? clock64() : 0;
    long long t_mark = t_begin;
    for (int attempt = 0; attempt < 2; ++attempt) {
#define PP_TICK(slot) do { if (A.timing && tid == 0) { const long long now_ = clock64(); S.cyc[slot] += now_ - t_mark; t_mark = now_; } } while (0)
    // ---- initializePlanner (LP:84-96) ------------------------------------------------
    {
      ulonglong2* hz = reinterpret_cast<ulonglong2*>(w.hash);
      for (int k = tid; k < A.hash_init / 2; k += kT) hz[k] = make_ulonglong2(0ull, 0ull);
      for (int k = tid; k < NC / 32 + 2; k += kT) w.cmin[k] = kCostInf;
    }
    if (tid == 0) {
      S.n_nodes = 0; S.n_closed = 0; S.heap_size = 0; S.n_exp = 0; S.status = -1; S.dup_seen = 0;
      S.n_cand_total = 0; S.n_coll_total = 0; S.goal_hit = 0; S.path_len = 0; S.truncated = 0; S.done = 0;
      S.fast = (attempt == 0 && A.fast_frontier) ? 1 : 0;
      S.hash_cur = A.hash_init;
      S.fx_slow = 0;
      S.tie_seen = 0; S.top_pos = 0; S.top_id = 0;
      for (int c = 0; c < 8; ++c) S.cyc[c] = 0;
      if (kBanked) S.gq = grid_at(A.g, A.bank, qq.reserved);
    }
    __syncthreads();
    if (warp == 0) {
      const double k0[6] = {0.0, 0.0, 0.0, 0.0, 0.0, Q.start_speed};
      open_single_warp0<kAllSmem>(w, heap, &S, A.hash_init, k0, 0.0, 9999999999.0, -1, lane);
    }
    __syncthreads();

    PP_TICK(0);
    // ---- while(true) of LP:31 ---------------------------------------------------------
    for (;;) {
      // caps replace the ros::Time guard of LP:33-39
      if (S.n_exp >= P.max_expansions || S.n_nodes >= P.max_nodes) { if (tid == 0) S.status = PP_PLAN_CAP; break; }
      if (S.heap_size == 0) { if (tid == 0) S.status = PP_PLAN_EXHAUSTED; break; }  // top() of an empty queue
      // the fingerprint set stays at most half full: rebuilt four times as large (from the stored fingerprints)
      // before an expansion could push it past that.  Most plans (a few thousand nodes) never leave the initial size.
      if (2 * (S.n_nodes + A.max_children + 2) > S.hash_cur && S.hash_cur < A.hash_size) {
        int grown = S.hash_cur;
        while (grown < A.hash_size && 2 * (S.n_nodes + A.max_children + 2) > grown) grown *= 4;
        if (grown > A.hash_size) grown = A.hash_size;
        const int nn = S.n_nodes;
        __syncthreads();
        ulonglong2* hz = reinterpret_cast<ulonglong2*>(w.hash);
        for (int k = tid; k < grown / 2; k += kT) hz[k] = make_ulonglong2(0ull, 0ull);
        __syncthreads();
        for (int id = tid; id < nn; id += kT) hash_insert(w, grown, w.fp[id]);
        if (tid == 0) S.hash_cur = grown;
        __syncthreads();
      }
      const int hash_cur = S.hash_cur;
      const int top_id = frontier_top<kAllSmem>(heap, w, &S, s_fmin);   // LP:40
      // current_node's fields are key fields, identical for every copy of the node
      const double ncx = w.cx[top_id], ncy = w.cy[top_id], npx = w.px[top_id], npy = w.py[top_id];
      const double nhd = w.hd[top_id], nvel = w.vel[top_id];
      __syncthreads();
      const int closed_id = close_top<kAllSmem>(w, heap, &S, top_id, s_red);                    // LP:41
      PP_TICK(1);
      if (tid == 0) {
        if (exp_out && S.n_exp < A.cap_exp) exp_out[S.n_exp] = closed_id;
        S.n_exp = S.n_exp + 1;
        S.goal_hit = 0;
      }
      __syncthreads();

      // ---- calcPossibleVelocities (LP:191-216) / calcPossibleYaws (LP:218-232) -----------------
      // every thread derives the same counts; no serial section
      const double time_step = dm::div(P.time_step_ms, 1000.0);
      const double v_hi = dm::add(nvel, dm::mul(Q.max_accel, time_step));
      const double v_lo = dm::sub(nvel, dm::mul(Q.max_accel, time_step));
      int n_vraw;
      if (!dm::trunc_to_int(dm::div(dm::sub(v_hi, v_lo), P.velocity_res), &n_vraw)) n_vraw = 0;
      bool unsupported = n_vraw > kMaxVelRaw;   // (the host has refused such queries already: PP_ERR_UNSUPPORTED)
      if (unsupported) n_vraw = 0;
      int n_vel = 0;
      for (int i = 0; i < n_vraw; ++i) {
        const double v = dm::add(v_lo, dm::mul(static_cast<double>(i), P.velocity_res));
        const bool skip = (v == 0 && Q.gs != 0) || v > Q.max_speed || v < 0;
        n_vel += skip ? 0 : 1;
      }
      // `for (int i = 0; i < double_n; i++)`: ceil of the double bound (the 26 / 27 trap)
      const double ylo = dm::sub(nhd, P.heading_diff);
      const double yn = dm::div(dm::sub(dm::add(nhd, P.heading_diff), ylo), P.heading_res);
      int n_yaw = 0;
      if (yn > 0) n_yaw = yn >= 4096.0 ? 4096 : __double2int_ru(yn);
      if (static_cast<long long>(n_vel) * n_yaw + S.n_nodes + 2 > A.NC) unsupported = true;
      const int n_cand = unsupported ? 0 : n_vel * n_yaw;
      const int base_nodes = S.n_nodes;
      const int hs_base = S.heap_size;
      int appended = 0;  // block-uniform running count
      int collided_local = 0;
      bool goal_stop = false;

      for (int c0 = 0; c0 == 0 || c0 < n_cand; c0 += kT) {
        // ---- checkForGoal (LP:269-287) over interpolatePoints(child, parent) (LP:301-316):
        // done by the last warp while the other warps already evaluate the children
        if (c0 == 0 && warp == kWarps - 1) {
          const double dx = dm::sub(ncx, npx), dy = dm::sub(ncy, npy);
          const double angle = dm::atan2_(dy, dx);
          const double dist = __dsqrt_rn(dm::add(dm::mul(dx, dx), dm::mul(dy, dy)));
          int n_pts;
          if (!dm::trunc_to_int(dm::mul(dm::div(dist, P.search_res), 2.0), &n_pts)) n_pts = 0;
          double sa, ca;
          dm::sincos(angle, &sa, &ca);
          // A18: P.goal_test_int_abs selects the `int abs(int)` reading of LP:276-283 (differences truncated
          // toward zero by the implicit conversion; outside +-2^30 = "too far", like the oracle's trunc_int)
          const bool int_abs = P.goal_test_int_abs != 0;
          const bool speed_ok = goal_tol_ok(dm::sub(nvel, Q.gs), P.goal_speed_tolerance, int_abs);
          for (int base = 0; base < n_pts; base += 32) {
            const int i = base + lane;
            bool hit = false;
            double x = 0, y = 0;
            if (i < n_pts) {
              const double t = dm::mul(dm::div(static_cast<double>(i), static_cast<double>(n_pts)), dist);
              x = dm::add(ncx, dm::mul(t, ca));
              y = dm::add(ncy, dm::mul(t, sa));
              hit = speed_ok && goal_tol_ok(dm::sub(x, Q.gx), P.goal_pos_tolerance, int_abs) &&
                    goal_tol_ok(dm::sub(y, Q.gy), P.goal_pos_tolerance, int_abs);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
              const int src = __ffs(m) - 1;  // first point in loop order
              if (lane == src) {
                // interpolateToGoalPoint (LP:289-299)
                S.goal_node[0] = x; S.goal_node[1] = y; S.goal_node[2] = ncx; S.goal_node[3] = ncy;
                S.goal_node[4] = dm::atan2_(dm::sub(y, ncy), dm::sub(x, ncx));
                S.goal_node[5] = Q.gs;
                S.goal_hit = 1;
              }
              break;
            }
          }
        }
        // ---- getNeighbors (LP:146-169): one thread per (velocity, yaw) child ------------------
        const int c = c0 + tid;
        bool alive = false, collided = false;
        int my_visit = -1;
        double k[6] = {0, 0, 0, 0, 0, 0};
        double g = 0, cost = 0;
        unsigned long long fp = 0;
        bool suspicious = false;
        if (c < n_cand) {
          const int vi = c / n_yaw, yi = c - vi * n_yaw;
          double vel = 0;
          int seen = 0;
          for (int i = 0; i < n_vraw; ++i) {   // the vi-th velocity that passes the filter of LP:205-212
            const double v = dm::add(v_lo, dm::mul(static_cast<double>(i), P.velocity_res));
            const bool skip = (v == 0 && Q.gs != 0) || v > Q.max_speed || v < 0;
            if (!skip) { if (seen == vi) vel = v; seen++; }
          }
          const double yaw = dm::add(ylo, dm::mul(static_cast<double>(yi), P.heading_res));
          const double sum_v = dm::add(nvel, vel);                                     // LP:155
          const double step = dm::div(dm::mul(sum_v, P.time_step_ms), 1000.0);
          double sy, cyaw;
          dm::sincos(yaw, &sy, &cyaw);
          const double x = dm::add(ncx, dm::mul(step, cyaw));                          // LP:156
          const double y = dm::add(ncy, dm::mul(step, sy));                            // LP:157
          // (kBanked is a template parameter on purpose: the plain instantiation keeps its grid pointers in
          // the constant bank; reading them through a run-time select cost 25 % on ORIENTED single plans)
          if (kBanked ? collide_pose_thread(S.gq, A.cp, x, y, yaw) : collide_pose_thread(A.g, A.cp, x, y, yaw)) {   // LP:159
            collided = true;
          } else {
            alive = true;
            k[0] = x; k[1] = y; k[2] = ncx; k[3] = ncy; k[4] = yaw; k[5] = vel;
            const double ddx = dm::sub(x, ncx), ddy = dm::sub(y, ncy);
            g = __dsqrt_rn(dm::add(dm::mul(ddx, ddx), dm::mul(ddy, ddy)));            // LP:162 (+ LP:139-144)
            cost = dm::add(g, calc_h(P, Q, KQ, x, y, yaw, sy, cyaw, vel));             // LP:163
            fp = key_fingerprint(k);
            suspicious = hash_contains(w, hash_cur, fp);
          }
        }
        // markVisited (LP:55 -> LP:318-326) for every child getNeighbors returned, bounds-guarded
        if (A.visited && c < n_cand && !collided) {
          int vx, vy;
          if (dm::trunc_to_int(dm::sub(static_cast<double>(A.g.W / 2), dm::div(k[1], A.g.res)), &vx) &&
              dm::trunc_to_int(dm::sub(static_cast<double>(A.g.H / 2), dm::div(k[0], A.g.res)), &vy) && vx >= 0 &&
              vx < A.g.W && vy >= 0 && vy < A.g.H)
            my_visit = vx * A.g.H + vy;
        }
        if (c0 == 0) {
          __syncthreads();
          PP_TICK(2);
          if (S.goal_hit) { goal_stop = true; break; }   // LP:43: the children are never looked at
        }
        if (unsupported) break;
        if (my_visit >= 0) A.visited[my_visit] = 1;   // only reached when the goal test did not end the plan
        // LP:56-72.  A fingerprint never seen before cannot equal any node: open it.  A seen
        // fingerprint takes the exact scan (rare).
        for (;;) {
          if (!__syncthreads_or(suspicious)) break;          // (the common case: one barrier instead of two + a reduction)
          const int owner = block_min(suspicious ? tid : kIntMax, s_red);
          if (owner == kIntMax) break;
          if (tid == owner) {
            for (int cix = 0; cix < 6; ++cix) S.scratch_key[cix] = k[cix];
          }
          __syncthreads();
          double kk[6];
          for (int cix = 0; cix < 6; ++cix) kk[cix] = S.scratch_key[cix];
          int fo, ac;
          scan_exact(w, S.n_nodes, kk, &fo, &ac, s_red);
          if (tid == owner) {
            if (ac) {
              alive = false;                               // in closed: dropped (LP:58)
            } else if (fo != kIntMax) {                    // in open: LP:66-70
              if (g < w.g[fo]) { w.g[fo] = g; w.cost[fo] = cost; }  // overwrite the stored copy in place
              S.dup_seen = 1;                              // ... and a duplicate is pushed anyway
            }
            suspicious = false;
          }
          __syncthreads();
        }
        // order-preserving compaction of the survivors of this chunk
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) s_scan[warp] = __popc(bal);
        collided_local += __syncthreads_count(collided);      // (the barrier of the scan also counts the collisions)
        int before = 0, total = 0;
        for (int ww = 0; ww < kWarps; ++ww) {
          const int v = s_scan[ww];
          if (ww < warp) before += v;
          total += v;
        }
        const int rank = appended + before + __popc(bal & ((1u << lane) - 1u));
        if (alive) {
          const int id = base_nodes + rank;
          w.cx[id] = k[0]; w.cy[id] = k[1]; w.px[id] = k[2]; w.py[id] = k[3]; w.hd[id] = k[4]; w.vel[id] = k[5];
          w.g[id] = g; w.cost[id] = cost;
          w.closed_seq[id] = -1;
          w.rev_last[id] = top_id;
          w.fp[id] = fp;
          hash_insert(w, hash_cur, fp);
          if (S.fast) {   // openNode's frontier push: a plain append, position = heap size + rank
            HeapEnt e;
            e.cost = cost; e.id = id; e.pad = 0;
            hset<kAllSmem>(heap, hs_base + rank, e);
            atomicMin(&w.cmin[(hs_base + rank) >> 5], cost_bits(cost));
          }
        }
        appended += total;
        __syncthreads();
      }
      if (goal_stop) {
        if (warp == 0) {
          double k[6];
          for (int c = 0; c < 6; ++c) k[c] = S.goal_node[c];
          open_single_warp0<kAllSmem>(w, heap, &S, hash_cur, k, 0.0, 0.0, top_id, lane);     // LP:45
        }
        __syncthreads();
        if (tid == 0) S.status = PP_PLAN_FOUND;
        break;
      }
      if (unsupported) { if (tid == 0) S.status = PP_PLAN_BROKEN; break; }
      PP_TICK(3);
      // openNode for every survivor, in order: frontier pushes by warp 0 (LP:62 / LP:70).  The
      // costs are fetched 32 at a time (one L2 round trip) and handed out by shuffle.
      if (warp == 0) {
        int hsize = S.heap_size;
        if (S.fast) hsize += appended;
        else
        for (int r0 = 0; r0 < appended; r0 += 32) {
          const int mine = r0 + lane;
          const double my_cost = mine < appended ? w.cost[base_nodes + mine] : 0.0;
          const int cnt = min(32, appended - r0);
          for (int r = 0; r < cnt; ++r) {
            const double cst = __shfl_sync(0xffffffffu, my_cost, r);
            heap_sift_up_warp<kAllSmem>(heap, hsize, cst, base_nodes + r0 + r, lane);
            hsize++;
          }
        }
        if (lane == 0) {
          S.heap_size = hsize;
          S.n_nodes = base_nodes + appended;
          S.n_cand_total += n_cand;
          S.n_coll_total += collided_local;
        }
      }
      __syncthreads();
      PP_TICK(4);
      if (S.heap_size == 0) { if (tid == 0) S.status = PP_PLAN_EXHAUSTED; break; }
      {
        const int top2 = frontier_top<kAllSmem>(heap, w, &S, s_fmin);
        __syncthreads();
        close_top<kAllSmem>(w, heap, &S, top2, s_red);                                          // LP:80
      }
      PP_TICK(5);
    }
    __syncthreads();
    if (!(S.fast && S.tie_seen)) break;   // the fast attempt is exact unless a pop saw a tied minimum
    if (A.visited) {                      // the re-run starts from a clean debug bitmap
      const size_t cells = static_cast<size_t>(A.g.W) * A.g.H;
      for (size_t c = tid; c < cells; c += kT) A.visited[c] = 0;
    }
    __syncthreads();
    }  // attempt

    PP_TICK(7);
    // ---- outputs -----------------------------------------------------------------------
    const int n_nodes = S.n_nodes;
    double* path = nullptr;
    int* path_ids = nullptr;
    double* prof = nullptr;
    int n_path = 0, n_profile = 0;
    if (S.status == PP_PLAN_FOUND) {
      // calcPathMsg (LP:503-527) and calcMPMessage (LP:542-583).  The walk is sequential; each
      // link search is ONE block-wide pass over the child points that yields three reductions:
      //   - first match at or after the scan position   (the link the reference follows)
      //   - first match overall                         (path_node_ids: LP:511-526 rule)
      //   - last match overall                          (the node calcMPMessage ends up with)
      // The frontier is dead now: its shared memory holds a copy of every child point, the
      // reversed path is staged in the workspace's frontier area.
      double* rev = w.heap_cost_g;   // 2*NC doubles: room for NC points
      int* rev_first = w.heap_id_g;  // slot of the first match per reversed pose
      int* rev_last = w.rev_last;    // slot of the last match per reversed pose
      const int rev_cap = NC;
      int len = 1;
      bool broken = false;
      // ---- pointer walk (the common case) --------------------------------------------------
      // Every node remembers the slot that was being expanded when it was opened (w.rev_last
      // during the search).  Following those slots from the terminal node gives a chain whose
      // consecutive child points are exactly parent_point links.  The reference walks by POINT
      // equality (LP:509-527) and resolves poses to nodes by point equality (LP:554-575); the two
      // agree whenever every pose of the chain is the child point of exactly one node, which one
      // parallel pass over the tree verifies.  Otherwise (a goal hit at i = 0 repeats a pose, a
      // duplicate was opened, ...) the exact scan walk below runs instead.
      bool fast_walk = false;
      {
        int* sp = reinterpret_cast<int*>(smem_raw);                  // parent slots in the dead frontier's smem
        const bool sp_fits = n_nodes <= A.heap_smem_entries * 4;
        if (sp_fits) for (int id = tid; id < n_nodes; id += kT) sp[id] = w.rev_last[id];
        for (int k = tid; k < kFxTable; k += kT) s_ptab[k] = 0;
        __syncthreads();
        if (tid == 0) {
          int id = n_nodes - 1, n = 0, bad = 0;                       // the terminal node was opened last (LP:45)
          for (;;) {
            if (n >= kFxMaxPoses || n >= rev_cap) { bad = 1; break; }
            rev_first[n++] = id;
            if (id == 0) break;                                       // the start node: child (0, 0), LP:509
            id = sp_fits ? sp[id] : w.rev_last[id];
            if (id < 0) { bad = 1; break; }
          }
          S.fx_len = n; S.fx_bad = bad;
        }
        __syncthreads();
        if (!S.fx_bad) {
          const int n = S.fx_len;
          for (int k = tid; k < n; k += kT) {
            const int id = rev_first[k];
            const double x = w.cx[id], y = w.cy[id];
            rev[2 * k] = x; rev[2 * k + 1] = y;
            unsigned short* slot = &s_ptab[point_slot(x, y, kFxTable - 1)];
            if (atomicCAS(slot, static_cast<unsigned short>(0), static_cast<unsigned short>(k + 1)) != 0) *slot = kFxShared;
          }
          __syncthreads();
          // One shared point is part of the common case: a goal hit at i = 0 (LP:276-283) makes the terminal
          // node T a copy of the popped node N's child point, so the chain starts T, N with identical poses and
          // that point has two owners.  The reference then walks T (open list, scanned first) -> N (closed list)
          // -> N's parent, i.e. exactly the chain; first match = T and last match = N for both poses.  Any OTHER
          // node on that point, or any other shared pose, still sends the query to the exact scan walk.
          const bool goal_dup = n >= 2 && rev[0] == rev[2] && rev[1] == rev[3];
          const int id_t = rev_first[0], id_n = n >= 2 ? rev_first[1] : -1;
          // (a goal hit on the start node itself ends the reference's walk before it begins, LP:509: scan walk)
          int bad = (goal_dup && rev[0] == 0.0 && rev[1] == 0.0) ? 1 : 0;
          for (int id = tid; id < n_nodes; id += kT) {
            const double x = w.cx[id], y = w.cy[id];
            // equal points hash alike, so only the pose(s) of this node's slot can share its child point
            const unsigned short e = s_ptab[point_slot(x, y, kFxTable - 1)];
            if (e == 0) continue;
            if (e != kFxShared) {
              const int k = e - 1;
              if (rev_first[k] != id && rev[2 * k] == x && rev[2 * k + 1] == y) bad = 1;
            } else {
              for (int k = 0; k < n; ++k)
                if (rev_first[k] != id && rev[2 * k] == x && rev[2 * k + 1] == y &&
                    !(goal_dup && k < 2 && (id == id_t || id == id_n)))
                  bad = 1;
            }
          }
          // consecutive poses must differ as well (LP:560 `else if`) -- implied by uniqueness, checked anyway
          for (int k = tid + (goal_dup ? 1 : 0); k + 1 < n; k += kT)
            if (rev[2 * k] == rev[2 * k + 2] && rev[2 * k + 1] == rev[2 * k + 3]) bad = 1;
          bad = block_max(bad, s_red);
          if (!bad) {
            fast_walk = true;
            len = n;
            if (!goal_dup) {
              rev_last = rev_first;
            } else {                                   // first / last match per pose as the scans would find them
              for (int k = tid; k < n; k += kT) rev_last[k] = k < 2 ? id_n : rev_first[k];
              __syncthreads();
              if (tid == 0) rev_first[1] = id_t;
            }
          }
        }
        __syncthreads();
      }
      if (!fast_walk) {
      if (tid == 0) S.fx_slow = S.fx_bad ? 1 : 2;
      // restore the slow path's scratch: w.rev_last is plain scratch from here on
      // a 32-bit hash of every child point fits the (dead) frontier's shared memory even in
      // batch mode; the scans below read it instead of the 16-byte points in L2 and verify
      // the rare hash hits against the real coordinates
      const uint32_t* sh32 = nullptr;
      if (n_nodes <= A.heap_smem_entries * 4) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
        for (int id = tid; id < n_nodes; id += kT) dst[id] = point_hash32(w.cx[id], w.cy[id]);
        sh32 = dst;
      }
      __syncthreads();
      // scan for child == (x, y): *k_from = min key >= from, *k_first = min key, *k_last = max key
      auto scan3 = [&](double x, double y, int from, int* k_from, int* k_first, int* k_last) {
        int bf = kIntMax, b0 = kIntMax, bl = -1;
        const uint32_t want = point_hash32(x, y);
        auto consider = [&](int id) {                   // hash hit (or no hashes): the real coordinates decide
          if (w.cx[id] == x && w.cy[id] == y) {
            const int cs = w.closed_seq[id];
            const int key = cs < 0 ? id : n_nodes + cs;
            b0 = min(b0, key);
            bl = max(bl, key);
            if (key >= from) bf = min(bf, key);
          }
        };
        // four hashes per 128-bit shared load -- only where registers are plentiful: in the 80-register batch
        // instantiation the wider loop spills and costs 4 % overall (measured), so it keeps the scalar scan
        if (kBatchCtas == 0 && sh32) {
          const uint4* sh4 = reinterpret_cast<const uint4*>(sh32);
          for (int id = 4 * tid; id < n_nodes; id += 4 * kT) {
            const uint4 v = sh4[id >> 2];
            if (v.x == want) consider(id);
            if (v.y == want && id + 1 < n_nodes) consider(id + 1);
            if (v.z == want && id + 2 < n_nodes) consider(id + 2);
            if (v.w == want && id + 3 < n_nodes) consider(id + 3);
          }
        } else if (sh32) {
          for (int id = tid; id < n_nodes; id += kT)
            if (sh32[id] == want) consider(id);
        } else {
          for (int id = tid; id < n_nodes; id += kT) consider(id);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          bf = min(bf, __shfl_xor_sync(0xffffffffu, bf, off));
          b0 = min(b0, __shfl_xor_sync(0xffffffffu, b0, off));
          bl = max(bl, __shfl_xor_sync(0xffffffffu, bl, off));
        }
        if (lane == 0) { s_red[warp] = bf; s_scan[warp] = b0; s_fpos[warp] = bl; }
        __syncthreads();
        bf = s_red[0]; b0 = s_scan[0]; bl = s_fpos[0];
        for (int ww = 1; ww < kWarps; ++ww) { bf = min(bf, s_red[ww]); b0 = min(b0, s_scan[ww]); bl = max(bl, s_fpos[ww]); }
        __syncthreads();
        *k_from = bf; *k_first = b0; *k_last = bl;
      };
      double cxp = S.goal_node[0], cyp = S.goal_node[1];
      if (tid == 0) { rev[0] = cxp; rev[1] = cyp; rev_first[0] = -1; rev_last[0] = -1; }
      while (!(cxp == 0 && cyp == 0)) {               // LP:509 -- checked between passes only
        int pos = 0;
        bool progressed = false;
        for (;;) {
          int key, kfirst, klast;
          scan3(cxp, cyp, pos, &key, &kfirst, &klast);
          if (tid == 0 && len - 1 < rev_cap) {          // ids of the pose currently being resolved
            rev_first[len - 1] = kfirst == kIntMax ? -1 : slot_of_key(w, n_nodes, kfirst);
            rev_last[len - 1] = klast < 0 ? -1 : slot_of_key(w, n_nodes, klast);
          }
          if (key == kIntMax) break;                   // end of this pass over open + closed
          const int id = slot_of_key(w, n_nodes, key);
          cxp = w.px[id]; cyp = w.py[id];
          if (len < rev_cap && tid == 0) { rev[2 * len] = cxp; rev[2 * len + 1] = cyp; rev_first[len] = -1; rev_last[len] = -1; }
          len++;
          progressed = true;
          pos = key + 1;
          if (len > n_nodes + 2) break;
        }
        if (!progressed || len > n_nodes + 2 || len >= rev_cap) { broken = true; break; }
      }
      if (!broken) {   // the terminal pose (0,0) has not been looked up by the walk if the last link closed a pass
        int key, kfirst, klast;
        scan3(cxp, cyp, 0, &key, &kfirst, &klast);
        if (tid == 0) {
          rev_first[len - 1] = kfirst == kIntMax ? -1 : slot_of_key(w, n_nodes, kfirst);
          rev_last[len - 1] = klast < 0 ? -1 : slot_of_key(w, n_nodes, klast);
        }
      }
      }  // !fast_walk
      __syncthreads();
      if (broken) {
        if (tid == 0) S.status = PP_PLAN_BROKEN;
      } else {
        n_path = len;
        if (A.rec_base) {
          unsigned char* rec = A.rec_base + static_cast<size_t>(q) * A.rec_stride + sizeof(pp_plan_record_header);
          path = reinterpret_cast<double*>(rec);
          prof = path + 2 * static_cast<size_t>(A.cap_path);
          path_ids = reinterpret_cast<int*>(prof + 3 * static_cast<size_t>(A.cap_path));
        } else {
          if (tid == 0) S.pool_off = static_cast<long long>(atomicAdd(A.pool_counter, static_cast<unsigned long long>(min(len, A.cap_path))));
          __syncthreads();
          path = A.out_path + 2 * S.pool_off;
          path_ids = A.out_path_ids + S.pool_off;
          prof = A.out_profile + S.pool_off;
        }
        const long long plane = A.rec_base ? static_cast<long long>(A.cap_path) : A.pool_cap;   // distance between the profile's three arrays
        for (int k = tid; k < len && k < A.cap_path; k += kT) {   // LP:530-538: reversed
          path[2 * k] = rev[2 * (len - 1 - k)];
          path[2 * k + 1] = rev[2 * (len - 1 - k) + 1];
          if (A.want_path_ids) path_ids[k] = rev_first[len - 1 - k];
        }
        if (len > A.cap_path && tid == 0) S.truncated = 1;
        // calcMPMessage: pose pair (i, i+1).  current_node = LAST node whose child == pose i;
        // next_node = LAST node whose child == pose i+1 -- but the reference's `else if` means a
        // node equal to pose i never counts for pose i+1, so identical consecutive poses leave
        // next_node at its default GraphNode(cur, next, 0, 0)
        n_profile = len >= 2 ? len - 2 : 0;
        for (int i = tid; i < n_profile && i < A.cap_path; i += kT) {
          const int ra = len - 1 - i, rb = len - 2 - i;
          const int id_a = rev_last[ra];
          const bool same = rev[2 * ra] == rev[2 * rb] && rev[2 * ra + 1] == rev[2 * rb + 1];
          const int id_b = same ? -1 : rev_last[rb];
          const double cur_heading = id_a >= 0 ? w.hd[id_a] : 0.0;
          const double next_heading = id_b >= 0 ? w.hd[id_b] : 0.0;
          const double next_velocity = id_b >= 0 ? w.vel[id_b] : 0.0;
          prof[i] = dm::div(dm::sub(next_heading, cur_heading), dm::div(P.time_step_ms, 1000.0));   // LP:576-577
          prof[plane + i] = P.time_step_ms;                                                          // LP:578
          prof[2 * plane + i] = next_velocity;                                                       // LP:579
        }
      }
    }
    __syncthreads();
    // tree digest over every slot's key
    unsigned long long dig = 0;
    for (int id = tid; id < n_nodes; id += kT) {
      const double k[6] = {w.cx[id], w.cy[id], w.px[id], w.py[id], w.hd[id], w.vel[id]};
      dig += node_digest(id, k);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dig += __shfl_xor_sync(0xffffffffu, dig, off);
    if (lane == 0) s_dig[warp] = dig;
    __syncthreads();
    if (tid == 0) {
      unsigned long long d = 0;
      for (int ww = 0; ww < kWarps; ++ww) d += s_dig[ww];
      PlanHdr hd;
      hd.status = S.status;
      hd.n_expansions = S.n_exp;
      hd.n_nodes = n_nodes;
      hd.n_closed = S.n_closed;
      hd.n_open = n_nodes - S.n_closed;
      hd.n_path = S.status == PP_PLAN_FOUND ? n_path : 0;
      hd.n_profile = S.status == PP_PLAN_FOUND ? n_profile : 0;
      hd.n_candidates = S.n_cand_total;
      hd.n_collisions = S.n_coll_total;
      hd.truncated = S.truncated;
      hd.reran = (!S.fast && A.fast_frontier) ? 1 : 0;
      hd.pad_ = S.fx_slow;
      hd.pool_off = S.status == PP_PLAN_FOUND ? S.pool_off : 0;
      hd.digest = d;
      if (A.timing) { const long long now_ = clock64(); S.cyc[6] += now_ - t_mark; S.cyc[7] = now_ - t_begin; }
      for (int c = 0; c < 8; ++c) hd.cyc[c] = S.cyc[c];
      A.hdr[q] = hd;
      if (A.rec_base) {
        pp_plan_record_header rh;
        rh.status = hd.status; rh.n_expansions = hd.n_expansions; rh.n_nodes = hd.n_nodes; rh.n_open = hd.n_open;
        rh.n_closed = hd.n_closed; rh.n_path = hd.n_path; rh.n_profile = hd.n_profile;
        rh.n_candidates = hd.n_candidates; rh.n_collisions = hd.n_collisions; rh.truncated = hd.truncated;
        rh.cap_path = A.cap_path; rh.reserved = 0; rh.tree_digest = d; rh.reserved2 = 0;
        *reinterpret_cast<pp_plan_record_header*>(A.rec_base + static_cast<size_t>(q) * A.rec_stride) = rh;
      }
    }
    if (A.rec_base) {
      // the unused tail of every array is zero, so that gathered records are comparable byte for byte
      unsigned char* rec = A.rec_base + static_cast<size_t>(q) * A.rec_stride + sizeof(pp_plan_record_header);
      double* r_path = reinterpret_cast<double*>(rec);
      double* r_prof = r_path + 2 * static_cast<size_t>(A.cap_path);
      int* r_ids = reinterpret_cast<int*>(r_prof + 3 * static_cast<size_t>(A.cap_path));
      const int np_ = S.status == PP_PLAN_FOUND ? min(n_path, A.cap_path) : 0;
      const int nm_ = S.status == PP_PLAN_FOUND ? min(n_profile, A.cap_path) : 0;
      for (int k = np_ + tid; k < A.cap_path; k += kT) { r_path[2 * k] = 0.0; r_path[2 * k + 1] = 0.0; r_ids[k] = -1; }
      for (int k = nm_ + tid; k < A.cap_path; k += kT) { r_prof[k] = 0.0; r_prof[A.cap_path + k] = 0.0; r_prof[2 * A.cap_path + k] = 0.0; }
      const size_t used = sizeof(pp_plan_record_header) + static_cast<size_t>(A.cap_path) * 44;
      for (size_t b = used + tid; b < A.rec_stride; b += kT) (A.rec_base + static_cast<size_t>(q) * A.rec_stride)[b] = 0;
    }
    __syncthreads();
  }
}

// node_parent (first node in scan order whose child_point == this node's parent_point,
// LP:511-526) for every slot of every query: one warp per target node.
__global__ void __launch_bounds__(256)
parent_ids_kernel(const unsigned char* ws_base, size_t ws_stride, int NC, int hash_size, const PlanHdr* hdr,
                  int n_queries, int* out_parent /* [n][NC] */) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long t = warp; t < static_cast<long long>(n_queries) * NC; t += n_warps) {
    const int q = static_cast<int>(t / NC), id = static_cast<int>(t % NC);
    const int n_nodes = hdr[q].n_nodes;
    if (id >= n_nodes) continue;
    const WS w = ws_view(const_cast<unsigned char*>(ws_base) + static_cast<size_t>(q) * ws_stride, NC, hash_size);
    const double x = w.px[id], y = w.py[id];
    int best = kIntMax, best_id = -1;
    for (int k = lane; k < n_nodes; k += 32) {
      if (w.cx[k] == x && w.cy[k] == y) {
        const int cs = w.closed_seq[k];
        const int key = cs < 0 ? k : n_nodes + cs;
        if (key < best) { best = key; best_id = k; }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, off), oi = __shfl_xor_sync(0xffffffffu, best_id, off);
      if (ob < best) { best = ob; best_id = oi; }
    }
    if (lane == 0) out_parent[static_cast<size_t>(q) * NC + id] = best_id;
  }
}

}  // namespace

}  // namespace pp

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct PlanWorkspace {
  unsigned char* ws = nullptr;
  size_t ws_bytes_total = 0;
  void* out = nullptr;          // hdr + path + path ids + profile + expansions
  size_t out_bytes = 0;
  pp_query* d_queries = nullptr;
  int queries_cap = 0;
  int* d_parent = nullptr;
  size_t parent_bytes = 0;
  int* d_counter = nullptr;
  unsigned char* h_out = nullptr;   // pinned mirror of the output block
  size_t h_out_bytes = 0;
};

namespace pp {

void plan_free(pp_handle* h) {
  if (!h->plan_ws) return;
  cudaFree(h->plan_ws->ws);
  cudaFree(h->plan_ws->out);
  cudaFree(h->plan_ws->d_queries);
  cudaFree(h->plan_ws->d_parent);
  cudaFree(h->plan_ws->d_counter);
  cudaFreeHost(h->plan_ws->h_out);
  delete h->plan_ws;
  h->plan_ws = nullptr;
}

static int grow(void** p, size_t* have, size_t want) {
  if (want <= *have) return PP_OK;
  cudaFree(*p);
  *p = nullptr;
  *have = 0;
  PP_CUDA(cudaMalloc(p, want));
  *have = want;
  return PP_OK;
}

// rec_dev != nullptr: record mode -- `rs` is ignored, every query leaves a fixed-stride record
// (pp_plan_record_stride(rec_cap_path) bytes) at rec_dev on the device, nothing is copied back and the call
// returns without waiting for the kernel.
int plan_batch(pp_handle* h, int n, const pp_query* qs, pp_result* rs, bool banked, unsigned char* rec_dev, int rec_cap_path) {
  if (!h->plan_ws) {
    h->plan_ws = new PlanWorkspace();
    PP_CUDA(cudaMalloc(&h->plan_ws->d_counter, sizeof(int)));
  }
  PlanWorkspace* W = h->plan_ws;
  cudaStream_t s = h->stream;
  const pp_params& P = h->P;
  PP_REQUIRE(P.max_nodes >= 1 && P.max_nodes <= (1 << 24), "pp_plan: max_nodes out of range");

  bool want_nodes = false, want_exp = false, want_path_ids = false;
  int cap_path = 2, cap_exp = 0;
  const bool record_mode = rec_dev != nullptr;
  if (record_mode) {
    PP_REQUIRE(rec_cap_path >= 2 && rec_cap_path <= (1 << 16), "pp_plan_batch_dev: cap_path must be in [2, 65536]");
    cap_path = rec_cap_path;
    want_path_ids = true;
  } else {
    for (int k = 0; k < n; ++k)
      PP_REQUIRE(rs[k].cap_nodes >= 0 && rs[k].cap_path >= 0 && rs[k].cap_expansions >= 0,
                 "pp_plan: negative capacity in a pp_result");
    for (int k = 0; k < n; ++k) {
      want_nodes |= rs[k].node_state || rs[k].node_parent || rs[k].node_closed_seq;
      want_exp |= rs[k].expansion_ids != nullptr;
      want_path_ids |= rs[k].path_node_ids != nullptr;
      cap_path = std::max(cap_path, rs[k].cap_path);
      if (rs[k].expansion_ids) cap_exp = std::max(cap_exp, rs[k].cap_expansions);
    }
    cap_path = std::min(cap_path, 1 << 16);
  }

  // one expansion can add up to kMaxCand children after the cap check, + the goal node
  // children one expansion can add after the cap check passed (+ the terminal node)
  long long max_children = 1;
  for (int k = 0; k < n; ++k) {
    const double nv = 2.0 * qs[k].max_accel * (P.time_step_ms / 1000.0) / P.velocity_res;
    const double ny = 2.0 * P.heading_diff / P.heading_res;
    if (!(nv < kMaxVelRaw - 2) || !(ny < 4094.0)) {   // (also catches NaN / inf parameters)
      set_last_error("pp_plan: more than 1024 velocities or 4096 yaws per expansion (max_accel * time_step / velocity_res, "
                     "heading_diff / heading_res): outside what the kernel implements");
      return PP_ERR_UNSUPPORTED;
    }
    const long long c = (static_cast<long long>(nv > 0 ? nv : 0) + 2) * (static_cast<long long>(ny > 0 ? ny : 0) + 2);
    max_children = std::max(max_children, c);
  }
  if (max_children > kMaxCand) {
    set_last_error("pp_plan: more than 8192 children per expansion (velocity_res / heading_res too fine)");
    return PP_ERR_UNSUPPORTED;
  }
  // multiple of 8 slots: every SoA sub-array of the workspace stays 16-byte aligned (the hash set is
  // cleared and the frontier entries are moved with 128-bit accesses)
  const int NC = ((P.max_nodes + static_cast<int>(max_children) + 64) + 7) & ~7;
  int hash_size = 1024;
  while (hash_size < 2 * NC) hash_size <<= 1;
  const size_t stride = ws_bytes(NC, hash_size);

  // Frontier placement: a lone query keeps the whole heap in shared memory (lowest latency);
  // a batch keeps only the top levels there so that several CTAs fit an SM and overlap each
  // other's serial heap phases.
  const int smem_cap_entries = 200 * 1024 / 16;
  int heap_entries = std::min(NC, smem_cap_entries);
  const bool wide = n > PP_PLAN_BATCH_CTAS * h->sm_count;          // more than one wave of the 6-CTA build
  if (wide) heap_entries = std::min(heap_entries, PP_PLAN_BATCH_HEAP_WIDE);
  else if (n >= 2 * h->sm_count) heap_entries = std::min(heap_entries, PP_PLAN_BATCH_HEAP);
  else if (n > h->sm_count) heap_entries = std::min(heap_entries, 6144);
  const size_t dyn_smem = 16 * static_cast<size_t>(heap_entries);
  const bool all_smem = heap_entries >= NC;
  int ctas_per_sm = 1;
  const int threads = n > h->sm_count ? 128 : 256;
  const bool batch = threads == 128;
  // the instantiation this call runs: frontier placement x CTA packing x (plain grid | grid bank)
  typedef void (*PlanKernel)(PlanArgs);
  static const PlanKernel kKernels[12] = {
      plan_kernel<false, 0, false>, plan_kernel<true, 0, false>,
      plan_kernel<false, PP_PLAN_BATCH_CTAS, false>, plan_kernel<true, PP_PLAN_BATCH_CTAS, false>,
      plan_kernel<false, PP_PLAN_BATCH_CTAS_WIDE, false>, plan_kernel<true, PP_PLAN_BATCH_CTAS_WIDE, false>,
      plan_kernel<false, 0, true>, plan_kernel<true, 0, true>,
      plan_kernel<false, PP_PLAN_BATCH_CTAS, true>, plan_kernel<true, PP_PLAN_BATCH_CTAS, true>,
      plan_kernel<false, PP_PLAN_BATCH_CTAS_WIDE, true>, plan_kernel<true, PP_PLAN_BATCH_CTAS_WIDE, true>};
  const PlanKernel kernel = kKernels[(all_smem ? 1 : 0) + (batch ? (wide ? 4 : 2) : 0) + (banked ? 6 : 0)];
  PP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, threads, dyn_smem));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int max_ctas = h->sm_count * ctas_per_sm;
  const int grid = std::min(n, max_ctas);
  const int n_ws = want_nodes ? n : grid;

  int rc = grow(reinterpret_cast<void**>(&W->ws), &W->ws_bytes_total, stride * static_cast<size_t>(n_ws));
  if (rc) return rc;
  // output block layout: headers | pool counter | path pool | id pool | 3 profile planes | expansions
  const long long pool_cap = record_mode ? 0 : static_cast<long long>(cap_path) * n;   // (records replace the pools)
  const size_t hdr_b = (sizeof(PlanHdr) * n + 255) & ~size_t(255);
  const size_t cnt_b = 256;
  const size_t path_b = (sizeof(double) * 2 * pool_cap + 255) & ~size_t(255);
  const size_t pid_b = (sizeof(int) * pool_cap + 255) & ~size_t(255);
  const size_t prof_b = (sizeof(double) * 3 * pool_cap + 255) & ~size_t(255);
  const size_t exp_b = want_exp ? ((sizeof(int) * static_cast<size_t>(cap_exp) * n + 255) & ~size_t(255)) : 0;
  const size_t out_total = hdr_b + cnt_b + path_b + pid_b + prof_b + exp_b;
  rc = grow(&W->out, &W->out_bytes, out_total);
  if (rc) return rc;
  if (!record_mode && out_total > W->h_out_bytes) {
    cudaFreeHost(W->h_out);
    W->h_out = nullptr; W->h_out_bytes = 0;
    PP_CUDA(cudaMallocHost(&W->h_out, out_total));
    W->h_out_bytes = out_total;
  }
  if (n > W->queries_cap) {
    cudaFree(W->d_queries);
    W->d_queries = nullptr;
    W->queries_cap = 0;
    PP_CUDA(cudaMalloc(&W->d_queries, sizeof(pp_query) * n));
    W->queries_cap = n;
  }
  PP_CUDA(cudaMemcpyAsync(W->d_queries, qs, sizeof(pp_query) * n, cudaMemcpyHostToDevice, s));
  PP_CUDA(cudaMemsetAsync(W->d_counter, 0, sizeof(int), s));

  unsigned char* ob = static_cast<unsigned char*>(W->out);
  PlanArgs A;
  A.g = banked ? h->bank_grid : h->grid;
  A.bank = banked ? h->bank : GridBank{};
  A.cp = collide_params(h);
  A.P = P;
  A.queries = W->d_queries;
  A.n_queries = n;
  A.NC = NC;
  A.hash_size = hash_size;
  // a found plan of the launch-file setting holds ~4-5 k nodes: 32 K entries (256 KB) cover 16 k nodes without a
  // rebuild (measured: a smaller start saves clearing but lengthens the probe chains: 16 K entries 10.9 ms, 32 K
  // 10.75 ms per 8192 queries), so raising max_nodes costs a batch nothing until a tree really grows that far;
  // a lone query starts at full size (no rebuild on the latency path)
  A.hash_init = n == 1 ? hash_size : std::min(hash_size, 32768);
  if (getenv("PP_PLAN_HASH_INIT")) A.hash_init = std::min(hash_size, std::max(1024, atoi(getenv("PP_PLAN_HASH_INIT"))));
  A.max_children = static_cast<int>(max_children);
  A.heap_smem_entries = heap_entries;
  A.ws_per_query = want_nodes ? 1 : 0;
  A.want_path_ids = want_path_ids ? 1 : 0;
  A.cap_path = cap_path;
  A.cap_exp = cap_exp;
  A.timing = getenv("PP_PLAN_TIMING") ? 1 : 0;
  A.fast_frontier = getenv("PP_PLAN_EXACT_HEAP") ? 0 : 1;
  A.ws_stride = stride;
  A.ws_base = W->ws;
  A.hdr = reinterpret_cast<PlanHdr*>(ob);
  A.pool_counter = reinterpret_cast<unsigned long long*>(ob + hdr_b);
  A.pool_cap = pool_cap;
  A.out_path = reinterpret_cast<double*>(ob + hdr_b + cnt_b);
  A.out_path_ids = reinterpret_cast<int*>(ob + hdr_b + cnt_b + path_b);
  A.out_profile = reinterpret_cast<double*>(ob + hdr_b + cnt_b + path_b + pid_b);
  A.out_exp = want_exp ? reinterpret_cast<int*>(ob + hdr_b + cnt_b + path_b + pid_b + prof_b) : nullptr;
  A.rec_base = rec_dev;
  A.rec_stride = record_mode ? pp_plan_record_stride(cap_path) : 0;
  PP_CUDA(cudaMemsetAsync(A.pool_counter, 0, sizeof(unsigned long long), s));
  A.work_counter = W->d_counter;
  A.visited = nullptr;
  if (n == 1 && !record_mode) {   // clearVisited (LP:251-262)
    const size_t cells = static_cast<size_t>(P.costmap_width) * P.costmap_height;
    if (!h->d_visited) PP_CUDA(cudaMalloc(&h->d_visited, cells));
    PP_CUDA(cudaMemsetAsync(h->d_visited, 0, cells, s));
    A.visited = h->d_visited;
  }

  PP_CUDA(cudaEventRecord(h->ev[8], s));
  kernel<<<grid, threads, dyn_smem, s>>>(A);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  bool want_parent = false;
  for (int k = 0; k < n && !record_mode; ++k) want_parent |= rs[k].node_parent != nullptr;
  if (want_parent) {
    rc = grow(reinterpret_cast<void**>(&W->d_parent), &W->parent_bytes, sizeof(int) * static_cast<size_t>(NC) * n);
    if (rc) return rc;
    parent_ids_kernel<<<h->sm_count * 8, 256, 0, s>>>(W->ws, stride, NC, hash_size, A.hdr, n, W->d_parent);
    h->launches++;
    PP_CUDA(cudaGetLastError());
  }
  PP_CUDA(cudaEventRecord(h->ev[9], s));
  h->plan_timing_pending = true;
  if (record_mode) return PP_OK;      // results stay on the device; the caller orders / waits on the handle's stream

  // headers + pool counter first, then only the used prefix of every pool
  unsigned char* hbw = W->h_out;
  PP_CUDA(cudaMemcpyAsync(hbw, W->out, hdr_b + cnt_b, cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  const size_t used = static_cast<size_t>(*reinterpret_cast<unsigned long long*>(hbw + hdr_b));
  if (used > 0) {
    const size_t o_path = hdr_b + cnt_b, o_pid = o_path + path_b, o_prof = o_pid + pid_b;
    PP_CUDA(cudaMemcpyAsync(hbw + o_path, ob + o_path, sizeof(double) * 2 * used, cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaMemcpyAsync(hbw + o_pid, ob + o_pid, sizeof(int) * used, cudaMemcpyDeviceToHost, s));
    for (int pl = 0; pl < 3; ++pl)
      PP_CUDA(cudaMemcpyAsync(hbw + o_prof + sizeof(double) * pool_cap * pl, ob + o_prof + sizeof(double) * pool_cap * pl,
                              sizeof(double) * used, cudaMemcpyDeviceToHost, s));
  }
  if (want_exp) {
    const size_t o_exp = hdr_b + cnt_b + path_b + pid_b + prof_b;
    PP_CUDA(cudaMemcpyAsync(hbw + o_exp, ob + o_exp, exp_b, cudaMemcpyDeviceToHost, s));
  }
  PP_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[8], h->ev[9]);
  h->phase_ms[PP_PHASE_PLAN] = ms;
  h->plan_timing_pending = false;

  const unsigned char* hb = W->h_out;
  const PlanHdr* hdr = reinterpret_cast<const PlanHdr*>(hb);
  if (getenv("PP_PLAN_TIMING")) {
    const char* names[8] = {"init", "close", "goal+eval", "commit", "push", "close2", "extract", "total"};
    for (int c = 0; c < 8; ++c) fprintf(stderr, "[pp_plan] query 0 %-10s %10lld cycles\n", names[c], hdr[0].cyc[c]);
    int reruns = 0;
    for (int k = 0; k < n; ++k) reruns += hdr[k].reran;
    fprintf(stderr, "[pp_plan] exact-heap re-runs (cost ties): %d of %d queries\n", reruns, n);
    int slow = 0, slow_dup = 0, found = 0;
    for (int k = 0; k < n; ++k) { slow += hdr[k].pad_ != 0; slow_dup += hdr[k].pad_ == 2; found += hdr[k].status == PP_PLAN_FOUND; }
    fprintf(stderr, "[pp_plan] paths extracted by the exact scan walk: %d of %d found plans (%d because a path pose is shared by two nodes)\n", slow, found, slow_dup);
  }
  const double* hpath = reinterpret_cast<const double*>(hb + hdr_b + cnt_b);
  const int* hpid = reinterpret_cast<const int*>(hb + hdr_b + cnt_b + path_b);
  const double* hprof = reinterpret_cast<const double*>(hb + hdr_b + cnt_b + path_b + pid_b);
  const int* hexp = want_exp ? reinterpret_cast<const int*>(hb + hdr_b + cnt_b + path_b + pid_b + prof_b) : nullptr;
  int ret = PP_OK;
  std::vector<double> soa;
  std::vector<int> iso;
  // scatter of the packed pools into the caller's per-query arrays: plain memcpys, split over a
  // few host threads for big batches (8192 queries = ~40 k small copies)
  std::atomic<int> cap_err{0};
  auto unpack = [&](int k_lo, int k_hi) {
    for (int k = k_lo; k < k_hi; ++k) {
      pp_result& r = rs[k];
      const PlanHdr& d = hdr[k];
      r.status = d.status;
      r.n_expansions = d.n_expansions;
      r.n_nodes = d.n_nodes;
      r.n_open = d.n_open;
      r.n_closed = d.n_closed;
      r.n_path = d.n_path;
      r.n_profile = d.n_profile;
      r.n_candidates = d.n_candidates;
      r.n_collisions = d.n_collisions;
      r.tree_digest = d.digest;
      const int np = std::min(d.n_path, std::min(r.cap_path, cap_path));
      if (d.n_path > r.cap_path && (r.path_xy || r.path_node_ids)) cap_err.store(1);
      if (r.path_xy) std::memcpy(r.path_xy, hpath + 2 * d.pool_off, sizeof(double) * 2 * np);
      if (r.path_node_ids) std::memcpy(r.path_node_ids, hpid + d.pool_off, sizeof(int) * np);
      const int nm = std::min(d.n_profile, std::min(r.cap_path, cap_path));
      const double* pr = hprof + d.pool_off;
      if (r.yaw_rates) std::memcpy(r.yaw_rates, pr, sizeof(double) * nm);
      if (r.durations) std::memcpy(r.durations, pr + pool_cap, sizeof(double) * nm);
      if (r.speeds) std::memcpy(r.speeds, pr + 2 * pool_cap, sizeof(double) * nm);
      if (r.expansion_ids) {
        const int ne = std::min(d.n_expansions, r.cap_expansions);
        if (d.n_expansions > r.cap_expansions) cap_err.store(1);
        std::memcpy(r.expansion_ids, hexp + static_cast<size_t>(k) * cap_exp, sizeof(int) * ne);
      }
    }
  };
  {
    const int n_thr = n >= 2048 ? 4 : 1;
    if (n_thr == 1) unpack(0, n);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < n_thr; ++t) th.emplace_back(unpack, static_cast<int>(static_cast<long long>(n) * t / n_thr),
                                                      static_cast<int>(static_cast<long long>(n) * (t + 1) / n_thr));
      for (auto& t : th) t.join();
    }
    if (cap_err.load()) ret = PP_ERR_CAPACITY;
  }
  for (int k = 0; k < n && want_nodes; ++k) {
    pp_result& r = rs[k];
    const PlanHdr& d = hdr[k];
    if (r.node_state || r.node_closed_seq || r.node_parent) {
      const int nn = std::min(d.n_nodes, r.cap_nodes);
      if (d.n_nodes > r.cap_nodes) ret = PP_ERR_CAPACITY;
      const unsigned char* wsq = W->ws + static_cast<size_t>(k) * stride;
      if (r.node_state) {
        soa.resize(static_cast<size_t>(8) * NC);
        PP_CUDA(cudaMemcpyAsync(soa.data(), wsq, sizeof(double) * 8 * NC, cudaMemcpyDeviceToHost, s));
        PP_CUDA(cudaStreamSynchronize(s));
        for (int id = 0; id < nn; ++id)
          for (int c = 0; c < 8; ++c) r.node_state[8 * static_cast<size_t>(id) + c] = soa[static_cast<size_t>(c) * NC + id];
      }
      if (r.node_closed_seq) {
        const size_t off = sizeof(double) * 8 * NC + sizeof(unsigned long long) * NC +
                           sizeof(unsigned long long) * hash_size + sizeof(double) * 2 * NC;
        PP_CUDA(cudaMemcpyAsync(r.node_closed_seq, wsq + off, sizeof(int) * nn, cudaMemcpyDeviceToHost, s));
      }
      if (r.node_parent)
        PP_CUDA(cudaMemcpyAsync(r.node_parent, W->d_parent + static_cast<size_t>(k) * NC, sizeof(int) * nn,
                                cudaMemcpyDeviceToHost, s));
      PP_CUDA(cudaStreamSynchronize(s));
    }
  }
  if (ret == PP_ERR_CAPACITY) set_last_error("pp_plan: a caller-provided output array is too small");
  return ret;
}

}  // namespace pp
