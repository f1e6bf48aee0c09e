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

// ---- REF_AABB (rows A10 / A11) --------------------------------------------------------------
// LP:471-472: base cell = int(W/2 - y/res), int(H/2 - x/res), float64, truncation toward zero.  The division
// is replaced by a multiplication with 1/res whenever that provably truncates to the same integer: both
// quotients are within 3 ulp of y/res, so for |result| < 2^20 the two differences lie within 1.2e-9 of each
// other and agree on the integer part unless the fast value is within 2^-20 of an integer -- those (and
// anything large or NaN) take the exact division.  The fast path is straight-line code (no conversion
// instructions, no branches), so the coordinates of several poses overlap in the pipelines.
constexpr int kAabbFar = -(1 << 30);      // "base cell" of a coordinate that cannot touch the map

// exact: LP:471-472 verbatim; coordinates outside +-2^30 cells (the conversion would be UB) never touch the map
__device__ __noinline__ int aabb_base_cell_exact(double half, double res, float v) {
  int out;
  if (!dm::trunc_to_int(dm::sub(half, dm::div(static_cast<double>(v), res)), &out)) return kAabbFar;
  return out;
}
// fast: returns false when the result needs the exact path
__device__ __forceinline__ bool aabb_base_cell_fast(double half, double inv_res, float v, int* out) {
  const double s = dm::fma(-static_cast<double>(v), inv_res, half);      // one rounding: at least as close as mul + sub
  // nearest integer of s without a conversion instruction: adding 1.5 * 2^52 leaves it in the low word of the sum
  const double t = dm::add(s, 6755399441055744.0);
  const double frac = dm::sub(s, dm::sub(t, 6755399441055744.0));       // in [-0.5, 0.5], exact
  // floor = nearest - (frac < 0); truncation toward zero adds one back for negative non-integers
  *out = __double2loint(t) + (__double2hiint(frac) >> 31) - (__double2hiint(s) >> 31);
  return fabs(frac) >= 9.5367431640625e-07 && fabs(s) < 1048576.0;      // (false for NaN)
}
__device__ __forceinline__ int aabb_base_cell(double half, double inv_res, double res, float v) {
  int out;
  if (aabb_base_cell_fast(half, inv_res, v, &out)) return out;
  return aabb_base_cell_exact(half, res, v);
}

// map cell of a pose + its 2-bit block summary: 0 = free (or outside the map), 1 = look the cell up, 3 = hit
struct AabbProbe {
  uint32_t cell;
  int code;
};
__device__ __forceinline__ AabbProbe aabb_probe_cells(const GridDev& g, const uint32_t* coarse, int bx, int by) {
  AabbProbe p;
  const int mx = bx + g.aabb_pad_x, my = by + g.aabb_pad_y;
  const bool inside = static_cast<unsigned>(mx) < static_cast<unsigned>(g.aabb_nx) &&
                      static_cast<unsigned>(my) < static_cast<unsigned>(g.aabb_ny);
  const int cx = inside ? mx >> g.aabb_cshift : 0, cy = inside ? my >> g.aabb_cshift : 0;
  const uint32_t w = coarse[cx * g.aabb_cwpr + (cy >> 4)];
  p.code = inside ? static_cast<int>((w >> ((cy & 15) * 2)) & 3u) : 0;
  p.cell = static_cast<uint32_t>(mx) * static_cast<uint32_t>(g.aabb_ny) + static_cast<uint32_t>(my);   // (< 2^31: grids <= 32768^2)
  return p;
}
__device__ __forceinline__ AabbProbe aabb_probe(const GridDev& g, const uint32_t* coarse, double half_w, double half_h,
                                                double inv_res, float x, float y) {
  return aabb_probe_cells(g, coarse, aabb_base_cell(half_w, inv_res, g.res, y), aabb_base_cell(half_h, inv_res, g.res, x));
}
__device__ __forceinline__ uint32_t aabb_resolve(const GridDev& g, const AabbProbe& p) {
  if (p.code == 1) return __ldg(&g.aabb_map[p.cell]) != 0 ? 1u : 0u;
  return static_cast<uint32_t>(p.code) >> 1;
}

// four poses at once (xs / ys = their coordinates): every stage is straight-line over the four, ONE branch for the
// (practically never taken) exact path, the cell look-ups as four independent predicated loads.  Returns the
// four flags in the bytes of a word.
__device__ __forceinline__ uint32_t aabb_quad(const GridDev& g, const uint32_t* coarse, double half_w, double half_h,
                                              double inv_res, const float xs[4], const float ys[4]) {
  int bx[4], by[4];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ok &= aabb_base_cell_fast(half_w, inv_res, ys[k], &bx[k]);
    ok &= aabb_base_cell_fast(half_h, inv_res, xs[k], &by[k]);
  }
  if (!ok) {            // some coordinate sits within 2^-20 of a cell border (or is huge / NaN): all eight exactly
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bx[k] = aabb_base_cell_exact(half_w, g.res, ys[k]);
      by[k] = aabb_base_cell_exact(half_h, g.res, xs[k]);
    }
  }
  AabbProbe p[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = aabb_probe_cells(g, coarse, bx[k], by[k]);
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint8_t m = p[k].code == 1 ? __ldg(&g.aabb_map[p[k].cell]) : static_cast<uint8_t>(p[k].code >> 1);
    out |= (m ? 1u : 0u) << (8 * k);
  }
  return out;
}

// plain variant (any alignment, any n): one pose per thread, the block summary read through L1
__global__ void __launch_bounds__(kThreads)
collide_aabb_kernel(GridDev g, const float* __restrict__ poses, uint8_t* __restrict__ flags, int64_t n) {
  const double half_w = static_cast<double>(g.W / 2), half_h = static_cast<double>(g.H / 2);
  const double inv_res = dm::div(1.0, g.res);
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const AabbProbe p = aabb_probe(g, g.aabb_coarse, half_w, half_h, inv_res, poses[3 * k], poses[3 * k + 1]);
    flags[k] = static_cast<uint8_t>(aabb_resolve(g, p));
  }
}

// 4 consecutive poses per thread straight from global memory (three 128-bit loads in flight per thread, poses
// 16-byte and flags 4-byte aligned), block summary through L1, one 32-bit store; the tail takes single poses
__global__ void __launch_bounds__(kThreads)
collide_aabb_quad_kernel(GridDev g, const float* __restrict__ poses, uint8_t* __restrict__ flags, int64_t n) {
  const double half_w = static_cast<double>(g.W / 2), half_h = static_cast<double>(g.H / 2);
  const double inv_res = dm::div(1.0, g.res);
  const int64_t n4 = n / 4;
  const float4* p4 = reinterpret_cast<const float4*>(poses);
  uint32_t* flags4 = reinterpret_cast<uint32_t*>(flags);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < n4; q += stride) {
    const float4 a = __ldg(&p4[3 * q]), b = __ldg(&p4[3 * q + 1]), c = __ldg(&p4[3 * q + 2]);
    const float xs[4] = {a.x, a.w, b.z, c.y}, ys[4] = {a.y, b.x, b.w, c.z};   // (x0 y0 w0 x1) (y1 w1 x2 y2) (w2 x3 y3 w3)
    flags4[q] = aabb_quad(g, g.aabb_coarse, half_w, half_h, inv_res, xs, ys);
  }
  for (int64_t k = n4 * 4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n; k += stride) {
    const AabbProbe p = aabb_probe(g, g.aabb_coarse, half_w, half_h, inv_res, poses[3 * k], poses[3 * k + 1]);
    flags[k] = static_cast<uint8_t>(aabb_resolve(g, p));
  }
}

// Streaming variant: the pose stream is the only HBM traffic that matters (12 B in, 1 B out per pose), so it is
// moved by the TMA unit.  One persistent CTA per SM = one PRODUCER warp + kAabbWarps CONSUMER warps around a ring of
// kAabbStages shared-memory buffers with a full / empty mbarrier pair each:
//   producer (one lane): wait empty[s] -> arm full[s] with the byte count -> ONE 1-D bulk copy (cp.async.bulk) of a
//                        2048-pose chunk (24 KB); completion is signalled on full[s] by the copy engine
//   consumer warp:       wait full[s] -> its 128 poses of the chunk, 4 consecutive poses per thread (three conflict-
//                        free LDS.128), block summary from shared memory, the rare cell look-ups from L2 as four
//                        independent predicated loads, one 32-bit store of 4 flags -> arrive on empty[s]
// No __syncthreads in the loop: warps drift up to kAabbStages - 1 chunks apart, which hides the L2 latency of the
// look-ups behind the other warps' arithmetic.
#ifndef PP_AABB_WARPS
#define PP_AABB_WARPS 16
#endif
#ifndef PP_AABB_CTAS
#define PP_AABB_CTAS 2
#endif
constexpr int kAabbWarps = PP_AABB_WARPS;             // consumer warps
constexpr int kAabbCtasPerSm = PP_AABB_CTAS;
constexpr int kAabbThreads = (kAabbWarps + 1) * 32;
constexpr int kAabbChunk = kAabbWarps * 32 * 4;       // poses per chunk (24 KB)
#ifndef PP_AABB_STAGES
#define PP_AABB_STAGES 3
#endif
constexpr int kAabbStages = PP_AABB_STAGES;
constexpr uint32_t kAabbChunkBytes = kAabbChunk * 12;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    if (spins == (1u << 26)) __trap();     // a copy that never lands is a bug, not something to wait out
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  }
}

__global__ void __launch_bounds__(kAabbThreads, kAabbCtasPerSm)
collide_aabb_stream_kernel(GridDev g, const float* __restrict__ poses, uint8_t* __restrict__ flags, int64_t n) {
  extern __shared__ __align__(128) unsigned char aabb_smem[];
  float* stage = reinterpret_cast<float*>(aabb_smem);
  uint32_t* s_coarse = reinterpret_cast<uint32_t*>(aabb_smem + static_cast<size_t>(kAabbStages) * kAabbChunkBytes);
  const int n_cw = g.aabb_cnbx * g.aabb_cwpr;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_coarse + ((n_cw + 3) & ~3));
  uint64_t* empty = full + kAabbStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_full = n / kAabbChunk;
  if (tid == 0) {
    for (int s = 0; s < kAabbStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kAabbWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                   // the barriers exist: the producer starts filling the ring at once ...
  if (warp != kAabbWarps) {                          // ... while the consumers fetch the block summary (17.7 KB at 4096^2)
    for (int k = tid; k < n_cw; k += kAabbWarps * 32) s_coarse[k] = __ldg(&g.aabb_coarse[k]);
    asm volatile("bar.sync 1, %0;" ::"n"(kAabbWarps * 32) : "memory");     // consumers only
  }
  const double half_w = static_cast<double>(g.W / 2), half_h = static_cast<double>(g.H / 2);
  const double inv_res = dm::div(1.0, g.res);
  if (warp == kAabbWarps) {
    // ---- producer ------------------------------------------------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t parity = 1;             // a fresh barrier passes a wait on the "previous" phase: the ring starts empty
      for (int64_t chunk = blockIdx.x; chunk < n_full; chunk += gridDim.x) {
        mbar_wait(&empty[s], parity);
        mbar_expect_tx(&full[s], kAabbChunkBytes);
        bulk_copy_g2s(stage + static_cast<size_t>(s) * kAabbChunk * 3, poses + chunk * kAabbChunk * 3, kAabbChunkBytes, &full[s]);
        if (++s == kAabbStages) { s = 0; parity ^= 1u; }
      }
    }
  } else {
    // ---- consumers -----------------------------------------------------------------------------------------
    uint32_t* flags4 = reinterpret_cast<uint32_t*>(flags);
    int s = 0;
    uint32_t parity = 0;
    for (int64_t chunk = blockIdx.x; chunk < n_full; chunk += gridDim.x) {
      mbar_wait(&full[s], parity);
      const float4* sp = reinterpret_cast<const float4*>(stage + static_cast<size_t>(s) * kAabbChunk * 3) + tid * 3;
      const float4 a = sp[0], b = sp[1], c = sp[2];      // (x0 y0 w0 x1) (y1 w1 x2 y2) (w2 x3 y3 w3)
      // release the buffer as soon as the poses are in registers: the (unused) operand makes the arrive wait for
      // all three loads of this lane, the shuffle-free __syncwarp for those of the other lanes' program order
      const uint32_t landed = __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(c.x);
      asm volatile("" ::"r"(landed) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      const float xs[4] = {a.x, a.w, b.z, c.y}, ys[4] = {a.y, b.x, b.w, c.z};
      const uint32_t out = aabb_quad(g, s_coarse, half_w, half_h, inv_res, xs, ys);
      flags4[chunk * (kAabbChunk / 4) + tid] = out;
      if (++s == kAabbStages) { s = 0; parity ^= 1u; }
    }
    // the last < 2048 poses: plain loads, spread over the CTAs
    for (int64_t k = n_full * kAabbChunk + blockIdx.x * static_cast<int64_t>(kAabbWarps * 32) + tid; k < n;
         k += static_cast<int64_t>(gridDim.x) * kAabbWarps * 32) {
      const AabbProbe p = aabb_probe(g, s_coarse, half_w, half_h, inv_res, poses[3 * k], poses[3 * k + 1]);
      flags[k] = static_cast<uint8_t>(aabb_resolve(g, p));
    }
  }
}

// compact host format (pp_collide_batch_q16): int16 triples -> float32 poses, one thread per pose
__global__ void __launch_bounds__(kThreads)
dequant_poses_kernel(const short* __restrict__ q, float* __restrict__ poses, int64_t n, pp_pose_quant k) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    poses[3 * i + 0] = __fadd_rn(k.x0, __fmul_rn(static_cast<float>(q[3 * i + 0]), k.dx));
    poses[3 * i + 1] = __fadd_rn(k.y0, __fmul_rn(static_cast<float>(q[3 * i + 1]), k.dy));
    poses[3 * i + 2] = __fmul_rn(static_cast<float>(q[3 * i + 2]), k.dyaw);
  }
}

// ---- flags <-> bits (the N > 1 result gather moves 1 bit per pose instead of 1 byte) ---------------------
// word k bit b = flags[32 k + b] != 0.  One thread per 16 flags (one 128-bit load), two threads per word.
__global__ void __launch_bounds__(kThreads)
pack_flags_kernel(const uint8_t* __restrict__ flags, uint32_t* __restrict__ bits, int64_t n) {
  const int64_t n_words = (n + 31) / 32;
  const int64_t total = n_words * 2;                   // (whole warps run every round: the shuffle below needs its neighbour)
  const int64_t total32 = (total + 31) & ~int64_t(31);
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total32;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint32_t v = 0;
    const int64_t base = t * 16;
    if (base + 16 <= n) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(flags + base));
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // any non-zero byte -> 1, then four 0/1 bytes -> one nibble: the partial products of the multiplication land
        // on distinct bit positions (no carries), bits 28..31 hold the result
        const uint32_t x = ws[q];
        const uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3) | (x >> 4) | (x >> 5) | (x >> 6) | (x >> 7)) & 0x01010101u;
        v |= ((nz * 0x10204080u) >> 28) << (4 * q);
      }
    } else {
      for (int b = 0; b < 16; ++b)
        if (base + b < n && flags[base + b]) v |= 1u << b;
    }
    const uint32_t hi = __shfl_down_sync(0xffffffffu, v, 1);
    if ((t & 1) == 0 && (t >> 1) < n_words) bits[t >> 1] = v | (hi << 16);
  }
}
__global__ void __launch_bounds__(kThreads)
unpack_flags_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ flags, int64_t n) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    flags[k] = (__ldg(&bits[k >> 5]) >> (k & 31)) & 1u;
}

__global__ void __launch_bounds__(kThreads)
fill_zero_kernel(uint8_t* __restrict__ flags, int64_t n) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    flags[k] = 0;
}

// One launch resolves a whole batch.  Every WARP works on its own: it draws blocks of 128 consecutive poses from a
// ticket counter and runs three phases over them, each with dense lanes, passing survivors along through its own
// slice of shared memory -- no CTA-wide barrier anywhere (with CTA-wide phases 36 % of the warp samples sat in
// __syncthreads waiting for the slowest warp of the phase):
//   1. lane per 4 poses : one bit of the 4x4-block map of the pose's yaw bin, dilated by the oriented footprint
//                        ("can the footprint, from anywhere in this block and at such a yaw, touch a block that
//                        holds an obstacle?") decides ~88 % of the poses.  All flags of the block are stored as 0
//                        here (one 32-bit store per lane); the later phases only store the 1s.  Survivors queue
//                        up until there are 32 of them.
//   2. lane per survivor : fp64 fixed-point lattice set-up, bounding box, oriented 4x4-block tests (make_job)
//                        -> free, colliding, or a NarrowJob
//   3. warp per job    : the tile rows under the box staged one job ahead (cp.async, one 128 B line per tile), a
//                        cheap pass along the lattice's perimeter, the hull test at cell resolution on the staged
//                        window, then the sample lattice
//   NU == 0 selects the run-time lattice shape.  broad_phase == 0 ("direct") disables every pruning test: each
//   pose that overlaps the grid runs the full lattice.
//   counters[3] += narrow-phase jobs of this launch; the low / high half of counters[1] = the ticket / the CTAs
//   finished: the last CTA to finish zeroes both, folds counters[3] into the running total counters[0] and
//   publishes it to pinned host memory (launches of one handle are serialised on its stream).
#ifndef PP_WIN_DEPTH
#define PP_WIN_DEPTH 1
#endif
constexpr int kWinDepth = PP_WIN_DEPTH;           // windows staged ahead of the one being tested
constexpr int kBlockPoses = 128;                  // poses per ticket: 4 consecutive poses per lane
constexpr int kQueueCap = 32 + kBlockPoses;       // phase 1 runs only while fewer than 32 survivors wait
constexpr size_t kOrientedWarpBytes = sizeof(uint32_t) * (kWinDepth + 1) * kWinWords + sizeof(NarrowJob) * 32 +
                                      sizeof(uint32_t) * 32 + sizeof(uint32_t) * kQueueCap;
static_assert(kOrientedWarpBytes % 16 == 0, "warp slices of shared memory stay 16-byte aligned");
constexpr size_t kOrientedSmem = kOrientedWarpBytes * kWarpsPerBlock;

#ifndef PP_LATTICE_S
#define PP_LATTICE_S 2
#endif
#ifndef PP_COLLIDE_MIN_BLOCKS
#define PP_COLLIDE_MIN_BLOCKS 3
#endif
// kBits: the output is ONE BIT per pose (bit b of word k = pose 32 k + b, what pp_pack_flags_dev makes of the byte
// flags): `flags` then points at ceil(n / 32) words.  Phase 1 zeroes the four words of its block, colliding poses
// OR their bit in.  For results that leave the device (the N > 1 gather) this saves the byte array and the pack pass.
template <int NU, int NV, int S, bool kBits>
__global__ void __launch_bounds__(kThreads, PP_COLLIDE_MIN_BLOCKS)
collide_oriented_kernel(GridDev g, Lattice lat, const float* __restrict__ poses, uint8_t* __restrict__ flags,
                        int64_t n, int broad_phase, unsigned long long* __restrict__ counters,
                        unsigned long long* __restrict__ host_pub, int first_of_call) {
  uint32_t* bits = reinterpret_cast<uint32_t*>(flags);
  // every warp's own slice of (dynamic) shared memory: ring of windows, jobs, their pose indices, survivor queue
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  unsigned char* mine_raw = s_raw + static_cast<size_t>(wib) * kOrientedWarpBytes;
  uint32_t (*s_win)[kWinWords] = reinterpret_cast<uint32_t (*)[kWinWords]>(mine_raw);
  NarrowJob* s_jobs = reinterpret_cast<NarrowJob*>(mine_raw + sizeof(uint32_t) * (kWinDepth + 1) * kWinWords);
  uint32_t* s_job_k = reinterpret_cast<uint32_t*>(s_jobs + 32);
  uint32_t* queue = s_job_k + 32;
  const uint32_t lt = (1u << lane) - 1u;
  const float inv_res_f = static_cast<float>(1.0 / g.res);
  const int reach = g.reach_cells;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(counters + 1);
  const bool vec = (reinterpret_cast<uintptr_t>(poses) & 15) == 0 && (reinterpret_cast<uintptr_t>(flags) & 3) == 0;
  unsigned long long local_narrow = 0;
  int nq = 0;
  bool more = true;
  for (;;) {
    // ---- phase 1: blocks of kBlockPoses poses until 32 survivors wait -----------------------------------
    while (more && nq < 32) {
      unsigned int blk = 0;
      if (lane == 0) blk = atomicAdd(ticket, 1u);
      blk = __shfl_sync(0xffffffffu, blk, 0);
      const int64_t k0 = static_cast<int64_t>(blk) * kBlockPoses + 4 * lane;
      if (static_cast<int64_t>(blk) * kBlockPoses >= n) { more = false; break; }
      float x[4], y[4], yaw[4];
      if (vec && k0 + 3 < n) {
        const float4* v = reinterpret_cast<const float4*>(poses + 3 * k0);
        const float4 a = __ldg(v), b = __ldg(v + 1), c = __ldg(v + 2);
        x[0] = a.x; y[0] = a.y; yaw[0] = a.z; x[1] = a.w; y[1] = b.x; yaw[1] = b.y;
        x[2] = b.z; y[2] = b.w; yaw[2] = c.x; x[3] = c.y; y[3] = c.z; yaw[3] = c.w;
        if (!kBits) *reinterpret_cast<uint32_t*>(flags + k0) = 0u;
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int64_t k = k0 + p;
          x[p] = y[p] = yaw[p] = 0.0f;
          if (k < n) { x[p] = poses[3 * k]; y[p] = poses[3 * k + 1]; yaw[p] = poses[3 * k + 2]; if (!kBits) flags[k] = 0; }
        }
      }
      if (kBits && lane < kBlockPoses / 32) {
        const int64_t wd = static_cast<int64_t>(blk) * (kBlockPoses / 32) + lane;
        if (wd * 32 < n) bits[wd] = 0u;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        bool survive = k0 + p < n;
        if (survive && broad_phase) {
          const float cif = static_cast<float>(g.W / 2) - y[p] * inv_res_f;
          const float cjf = static_cast<float>(g.H / 2) - x[p] * inv_res_f;
          if (fabsf(cif) < 1.0e6f && fabsf(cjf) < 1.0e6f && fabsf(yaw[p]) < 100.0f) {   // (NaN / absurd poses take the exact path)
            const int ic = __float2int_rd(cif), jc = __float2int_rd(cjf);
            if (ic + reach < 0 || ic - reach >= g.W || jc + reach < 0 || jc - reach >= g.H) {
              survive = false;                               // the footprint cannot reach the grid
            } else {
              const int bin = __float2int_rd(yaw[p] * (kNearBins / 3.14159265f)) & (kNearBins - 1);
              const int r4 = (ic + kHaloCells) >> 2, c4 = (jc + kHaloCells) >> 2;
              survive = (__ldg(&g.near4[(static_cast<size_t>(bin) * g.CI4 + r4) * g.CW4 + (c4 >> 6)]) >> (c4 & 63)) & 1ull;
            }
          }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, survive);
        if (survive) queue[nq + __popc(bal & lt)] = static_cast<uint32_t>(k0 + p);
        nq += __popc(bal);
      }
    }
    if (nq == 0) break;
    __syncwarp();                                            // queue entries and the zeroed flags are in place
    // ---- phase 2: the last (up to) 32 survivors, one per lane --------------------------------------
    const int take = nq < 32 ? nq : 32;
    nq -= take;
    int verdict = 0;
    NarrowJob job;
    uint32_t k = 0;
    if (lane < take) {
      k = queue[nq + lane];
      const float px = poses[3 * static_cast<int64_t>(k)], py = poses[3 * static_cast<int64_t>(k) + 1],
                  pyaw = poses[3 * static_cast<int64_t>(k) + 2];
      const PoseFx fx = pose_to_fixed(lat, g.W, g.H, g.res, static_cast<double>(px), static_cast<double>(py),
                                      static_cast<double>(pyaw));
      verdict = make_job(g, lat, fx, broad_phase, &job);
      if (verdict == 2) {                                    // decided by the coarse maps
        if (kBits) atomicOr(&bits[k >> 5], 1u << (k & 31u));
        else flags[k] = 1;
      }
    }
    const uint32_t jbal = __ballot_sync(0xffffffffu, verdict == 1);
    if (verdict == 1) {
      const int slot = __popc(jbal & lt);
      s_jobs[slot] = job;
      s_job_k[slot] = k;
    }
    const int nj = __popc(jbal);
    local_narrow += nj;
    __syncwarp();
    // ---- phase 3: the warp takes the jobs one after the other ----------------------------------------
    // The windows of the next kWinDepth jobs are on their way (asynchronous copies into a ring of windows)
    // while the lattice of the current one runs: one commit group per job, empty past the last job.
    bool mine = false;                                       // lane j keeps job j's verdict
#pragma unroll
    for (int j = 0; j < kWinDepth; ++j) {
      if (j < nj) stage_window_async(g, s_jobs[j], s_win[j], lane);
      async_commit();
    }
    for (int j = 0; j < nj; ++j) {
      if (j + kWinDepth < nj) stage_window_async(g, s_jobs[j + kWinDepth], s_win[(j + kWinDepth) % (kWinDepth + 1)], lane);
      async_commit();
      async_wait<kWinDepth>();
      __syncwarp();
      const NarrowJob& jb = s_jobs[j];                       // (fields are fetched where they are used)
      const uint32_t* win = s_win[j % (kWinDepth + 1)];
      bool hit;
      if (broad_phase) hit = window_collide_fixed<NU, NV, S>(lat, jb, win, lane);
      else if (NU > 0) hit = lattice_any_fixed<(NU > 0 ? NU : 1), (NU > 0 ? NV : 1), S>(jb, win, lane);
      else hit = lattice_any_generic(lat, jb, win, lane);
      __syncwarp();                                          // (this window is refilled by the copy issued next)
      if (lane == j) mine = hit;
    }
    if (lane < nj && mine) {
      const uint32_t k = s_job_k[lane];
      if (kBits) atomicOr(&bits[k >> 5], 1u << (k & 31u));
      else flags[k] = 1;
    }
    __syncwarp();                                            // s_jobs / s_job_k are rewritten by the next pass
  }
  if (lane == 0 && local_narrow) atomicAdd(counters + 3, local_narrow);
  __syncthreads();
  if (tid == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(counters + 1) + 1;
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      // the last CTA of the launch: everything is left ready for the next launch (no memset between launches), and
      // the job counts go straight to the pinned host words {total before this call, total now} -- a call may be
      // several launches; pp_last_narrow_count takes the difference once the stream is idle.
      __threadfence();
      const unsigned long long before = counters[0];
      const unsigned long long now = before + *reinterpret_cast<volatile unsigned long long*>(counters + 3);
      counters[0] = now; counters[3] = 0ull;
      *ticket = 0u; *done = 0u;
      if (first_of_call) host_pub[0] = before;
      host_pub[1] = now;
      __threadfence_system();
    }
  }
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
          need = make_job_light(g, cp.lat, pf, &job);
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
int collide_launch(pp_handle* h, int64_t n, const float* poses, uint8_t* flags, bool bits_out) {
  if (n == 0) return PP_OK;
  cudaStream_t s = h->stream;
  const int mode = h->P.collision_mode;
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(h->d_count);
  if (bits_out && mode != PP_COLLISION_ORIENTED) return PP_ERR_INVALID;   // (pp_collide_batch_bits_dev packs the other modes' flags)
  if (mode == PP_COLLISION_AS_SHIPPED) {
    fill_zero_kernel<<<grid_blocks(h, n, kThreads * 16, 8), kThreads, 0, s>>>(flags, n);
  } else if (mode == PP_COLLISION_REF_AABB) {
    const GridDev& g = h->grid;
    const size_t smem = static_cast<size_t>(kAabbStages) * kAabbChunkBytes +
                        sizeof(uint32_t) * ((static_cast<size_t>(g.aabb_cnbx) * g.aabb_cwpr + 3) & ~size_t(3)) + 16 * kAabbStages;
    const bool aligned = (reinterpret_cast<uintptr_t>(poses) & 15) == 0 && (reinterpret_cast<uintptr_t>(flags) & 3) == 0;
    static const int variant = getenv("PP_AABB_VARIANT") ? atoi(getenv("PP_AABB_VARIANT")) : 2;   // 0 plain, 1 quad, 2 TMA stream
    if (aligned && n >= 4 * kAabbChunk && variant == 2) {
      if (!h->aabb_stream_ready) {       // (per device: the attribute lives in the context)
        PP_CUDA(cudaFuncSetAttribute(collide_aabb_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        h->aabb_stream_ready = true;
      }
      const int64_t chunks = n / kAabbChunk;
      const int64_t cap = static_cast<int64_t>(kAabbCtasPerSm) * h->sm_count;   // persistent CTAs
      collide_aabb_stream_kernel<<<static_cast<int>(chunks < cap ? chunks : cap), kAabbThreads, smem, s>>>(g, poses, flags, n);
    } else if (aligned && variant >= 1) {
      collide_aabb_quad_kernel<<<grid_blocks(h, (n + 3) / 4, kThreads, 8), kThreads, 0, s>>>(g, poses, flags, n);
    } else {
      collide_aabb_kernel<<<grid_blocks(h, n, kThreads, 8), kThreads, 0, s>>>(g, poses, flags, n);
    }
  } else {
    // persistent: as many CTAs as fit (4 per SM), every warp draws 128-pose blocks from the ticket counter
    static const int bps = getenv("PP_COLLIDE_BLOCKS_PER_SM") ? atoi(getenv("PP_COLLIDE_BLOCKS_PER_SM")) : PP_COLLIDE_MIN_BLOCKS;
    constexpr int64_t kSlice = int64_t(1) << 31;     // pose indices travel as 32-bit words in shared memory
    if (!h->oriented_ready) {                        // (per device: the attribute lives in the context)
      const int sm = static_cast<int>(kOrientedSmem);
      PP_CUDA(cudaFuncSetAttribute(collide_oriented_kernel<48, 17, PP_LATTICE_S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      PP_CUDA(cudaFuncSetAttribute(collide_oriented_kernel<20, 8, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      PP_CUDA(cudaFuncSetAttribute(collide_oriented_kernel<0, 0, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      PP_CUDA(cudaFuncSetAttribute(collide_oriented_kernel<48, 17, PP_LATTICE_S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      PP_CUDA(cudaFuncSetAttribute(collide_oriented_kernel<20, 8, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      PP_CUDA(cudaFuncSetAttribute(collide_oriented_kernel<0, 0, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      h->oriented_ready = true;
    }
    for (int64_t lo = 0; lo < n; lo += kSlice) {
      const int64_t m = n - lo < kSlice ? n - lo : kSlice;
      const float* ps = poses + 3 * lo;
      uint8_t* fs = bits_out ? flags + lo / 8 : flags + lo;      // (slices start on a multiple of 32 poses)
      const int blocks = grid_blocks(h, m, kBlockPoses * kWarpsPerBlock, bps);
      unsigned long long* pub = reinterpret_cast<unsigned long long*>(h->h_count);   // (pinned: the device writes through UVA)
      const int first = h->narrow_first ? 1 : 0;
      h->narrow_first = false;
#define PP_LAUNCH_ORIENTED(NU, NV, S)                                                                                         \
  do {                                                                                                                        \
    if (bits_out) collide_oriented_kernel<NU, NV, S, true><<<blocks, kThreads, kOrientedSmem, s>>>(h->grid, h->lat, ps, fs, m, h->broad_phase, counter, pub, first); \
    else collide_oriented_kernel<NU, NV, S, false><<<blocks, kThreads, kOrientedSmem, s>>>(h->grid, h->lat, ps, fs, m, h->broad_phase, counter, pub, first);        \
  } while (0)
      if (h->lat.nu == 48 && h->lat.nv == 17) PP_LAUNCH_ORIENTED(48, 17, PP_LATTICE_S);   // 4.7 m x 1.586 m at 0.1 m cells
      else if (h->lat.nu == 20 && h->lat.nv == 8) PP_LAUNCH_ORIENTED(20, 8, 4);          // the same car at the launch file's 0.25 m cells
      else PP_LAUNCH_ORIENTED(0, 0, 1);
#undef PP_LAUNCH_ORIENTED
    }
  }
  h->launches++;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}

// A "call" may be several launches (chunks of the host pipeline, slices of a huge batch): the first ORIENTED launch
// after this publishes the running total it starts from, every launch publishes the total it ends with
// (collide_oriented_kernel's last CTA), and pp_last_narrow_count takes the difference.  No stream operation.
int collide_counter_reset(pp_handle* h) {
  h->narrow_first = true;
  return PP_OK;
}
int collide_counter_fetch(pp_handle* h) {
  h->last_narrow = -1;  // resolved from the pinned words once the stream is idle
  return PP_OK;
}

int collide_batch_dev(pp_handle* h, int64_t n, const float* poses, uint8_t* flags, bool bits_out) {
  if (n == 0) return PP_OK;
  cudaStream_t s = h->stream;
  const bool counted = h->P.collision_mode == PP_COLLISION_ORIENTED;   // only the ORIENTED kernel counts narrow-phase jobs
  int rc = counted ? collide_counter_reset(h) : PP_OK;
  if (rc) return rc;
  PP_CUDA(cudaEventRecord(h->ev[2], s));
  rc = collide_launch(h, n, poses, flags, bits_out);
  if (rc) return rc;
  PP_CUDA(cudaEventRecord(h->ev[3], s));
  rc = counted ? collide_counter_fetch(h) : PP_OK;
  if (rc) return rc;
  if (!counted) h->last_narrow = n;
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

int dequant_poses_dev(pp_handle* h, int64_t n, const int16_t* q, const pp_pose_quant& k, float* poses) {
  if (n == 0) return PP_OK;
  dequant_poses_kernel<<<grid_blocks(h, n, kThreads, 8), kThreads, 0, h->stream>>>(q, poses, n, k);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}

int pack_flags_dev(pp_handle* h, int64_t n, const uint8_t* flags, uint32_t* bits) {
  if (n == 0) return PP_OK;
  if (reinterpret_cast<uintptr_t>(flags) & 15) { set_last_error("pp_pack_flags_dev: flags must be 16-byte aligned"); return PP_ERR_INVALID; }
  pack_flags_kernel<<<grid_blocks(h, (n + 31) / 32 * 2, kThreads, 8), kThreads, 0, h->stream>>>(flags, bits, n);
  h->launches++;
  PP_CUDA(cudaGetLastError());
  return PP_OK;
}
int unpack_flags_dev(pp_handle* h, int64_t n, const uint32_t* bits, uint8_t* flags) {
  if (n == 0) return PP_OK;
  unpack_flags_kernel<<<grid_blocks(h, n, kThreads, 8), kThreads, 0, h->stream>>>(bits, flags, n);
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
