// pp_narrow.cuh -- warp-cooperative ORIENTED footprint test shared by the batch collision
// kernel (pp_collide.cu) and the sampling planner's edge kernel (pp_rrt.cu).
//
//   broad phase  (one thread):  fixed-point lattice set-up, bounding box, then
//       make_job        (batch kernel)   the oriented 4x4-block test: free / colliding / NarrowJob
//       make_job_light  (edges)          tile flags + 4x4 blocks under the bounding box: free / NarrowJob
//   narrow phase (one warp):    the three tile rows under the box staged into shared memory (each tile = one
//       coalesced 128 B line), then
//       window_collide_fixed (batch)     perimeter pass, cell-level hull test, the nu x nv sample lattice
//       window_box_collide   (edges)     exact box test, the sample lattice
#pragma once
#pragma once
#include "pp_common.cuh"

namespace pp {

// What the narrow phase needs to know about one pose, relative to the corner of the first
// tile under its bounding box (so all lattice coordinates are in [0, 96) cells).
struct __align__(16) NarrowJob {
  int32_t oi, oj, ui, uj, vi, vj;
  uint32_t tiles;   // (t0i * TJ + t0j) * 32: word offset of the first tile under the box in occ_tiles
  uint32_t box;     // ri0 | ri1 << 7 | rj0 << 14 | rj1 << 21 : lattice bbox relative to the tile corner
  uint32_t occ;     // bits 3a .. 3a+2: tile row t0i+a is under the box AND holds an obstacle in the box's columns
  // the footprint's two bands (footprint_bands) in window coordinates, hull over ONE cell row:  the columns the
  // samples of window row r can reach are  max(kT d + tlo, kS d + slo) .. min(kT d + thi, kS d + shi),  d = r + d0
  float kT, kS, tlo, thi, slo, shi, d0;
};                  // 64 bytes: four 128-bit shared-memory loads per job

// Broad phase for one pose (one thread): 0 = certainly collision-free, 2 = certainly colliding (nothing written to
// *job in either case), 1 = *job describes the pose for the narrow phase.
__device__ __forceinline__ int make_job(const GridDev& g, const Lattice& l, const PoseFx& p, int broad_phase,
                                        NarrowJob* job) {
  if (!p.valid) return 0;
  int i0, i1, j0, j1;
  lattice_bbox(l, p, &i0, &i1, &j0, &j1);
  if (i1 < 0 || i0 >= g.W || j1 < 0 || j0 >= g.H) return 0;      // box entirely off the grid
  const int t0i = (i0 + kHaloCells) >> 5, t0j = (j0 + kHaloCells) >> 5;
  const int nti = ((i1 + kHaloCells) >> 5) - t0i + 1;
  uint32_t occ;
  Bands B;
  if (broad_phase) {
    // the 4x4 blocks under the footprint decide most poses, and tell which tile rows hold an obstacle in the box
    B = footprint_bands(l, p);
    uint32_t rows;
    const int verdict = coarse4_oriented_test<kCoarse4Unroll>(g, B, p, i0, i1, j0, j1, &rows);
    if (verdict != 1) return verdict;
    occ = ((rows & 1u) ? 7u : 0u) | ((rows & 2u) ? 56u : 0u) | ((rows & 4u) ? 448u : 0u);
  } else {
    B.kT = B.kS = 0.0f; B.eT0 = B.eS0 = -kBandBig; B.eT1 = B.eS1 = kBandBig; B.core = 0;
    occ = (nti >= 3 ? 511u : (nti == 2 ? 63u : 7u));               // direct mode: every tile row under the box
  }
  // re-base the lattice origin from the anchor cell to the first tile's corner
  const int ci0 = t0i * kTile - kHaloCells, cj0 = t0j * kTile - kHaloCells;
  job->oi = p.oi + ((p.ai - ci0) << kFracBits);
  job->oj = p.oj + ((p.aj - cj0) << kFracBits);
  job->ui = p.ui; job->uj = p.uj; job->vi = p.vi; job->vj = p.vj;
  job->tiles = (static_cast<uint32_t>(t0i) * static_cast<uint32_t>(g.TJ) + static_cast<uint32_t>(t0j)) * kTile;
  job->box = static_cast<uint32_t>(i0 - ci0) | (static_cast<uint32_t>(i1 - ci0) << 7) |
             (static_cast<uint32_t>(j0 - cj0) << 14) | (static_cast<uint32_t>(j1 - cj0) << 21);
  job->occ = occ;
  {
    constexpr float q = 1.0f / static_cast<float>(1 << kFracBits);
    constexpr float kM = kBandMargin, kStrip = 1.0f + 2.0f * kM;
    const float oj = static_cast<float>(job->oj) * q;
    const float wT = kStrip * B.kT, wS = kStrip * B.kS;
    job->kT = B.kT; job->kS = B.kS;
    job->tlo = B.eT0 + fminf(0.0f, wT) + (oj - kM); job->thi = B.eT1 + fmaxf(0.0f, wT) + (oj + kM);
    job->slo = B.eS0 + fminf(0.0f, wS) + (oj - kM); job->shi = B.eS1 + fmaxf(0.0f, wS) + (oj + kM);
    job->d0 = -(static_cast<float>(job->oi) * q) - kM;
  }
  return 1;
}

// The LIGHT broad phase, for callers where a few warps wait on one pose at a time (the edges of the sampling
// planner, pp_collide_edges_dev): tile flags and the 4x4 blocks under the bounding box -- independent loads, a
// fraction of the oriented test's instructions.  It decides fewer poses, which suits those callers: their jobs are
// few and the latency of the set-up is what they pay for (config 4: 20.7 ms with this, 28.9-30.7 ms with make_job).
// false = certainly collision-free; the job's bands are left unset (window_box_collide does not use them).
__device__ __forceinline__ bool make_job_light(const GridDev& g, const Lattice& l, const PoseFx& p, NarrowJob* job) {
  if (!p.valid) return false;
  int i0, i1, j0, j1;
  lattice_bbox(l, p, &i0, &i1, &j0, &j1);
  if (i1 < 0 || i0 >= g.W || j1 < 0 || j0 >= g.H) return false;  // box entirely off the grid
  const int t0i = (i0 + kHaloCells) >> 5, t0j = (j0 + kHaloCells) >> 5;
  const int nti = ((i1 + kHaloCells) >> 5) - t0i + 1, ntj = ((j1 + kHaloCells) >> 5) - t0j + 1;
  uint32_t occ = 0;
  for (int a = 0; a < nti; ++a)
    for (int b = 0; b < ntj; ++b)
      occ |= (__ldg(&g.tile_any[static_cast<size_t>(t0i + a) * g.TJ + (t0j + b)]) ? 1u : 0u) << (a * 3 + b);
  if (occ == 0) return false;                                    // no obstacle in any tile under the box
  {
    const int r0 = (i0 + kHaloCells) >> 2, r1 = (i1 + kHaloCells) >> 2;
    const int c0 = (j0 + kHaloCells) >> 2, c1 = (j1 + kHaloCells) >> 2;
    const int w0 = c0 >> 6, sh = c0 & 63, nb = c1 - c0 + 1;      // nb <= 16 bits starting at bit c0
    const unsigned long long m = ((1ull << nb) - 1ull);
    unsigned long long any = 0;
    for (int r = r0; r <= r1; ++r) {
      const unsigned long long* row = g.coarse4 + 2 * (static_cast<size_t>(r) * g.CW4 + w0);   // "any" words of the pairs
      unsigned long long bits = __ldg(row) >> sh;
      if (sh + nb > 64) bits |= __ldg(row + 2) << (64 - sh);
      any |= bits & m;
    }
    if (!any) return false;                                      // none in the 4x4 blocks under it
  }
  const int ci0 = t0i * kTile - kHaloCells, cj0 = t0j * kTile - kHaloCells;
  job->oi = p.oi + ((p.ai - ci0) << kFracBits);
  job->oj = p.oj + ((p.aj - cj0) << kFracBits);
  job->ui = p.ui; job->uj = p.uj; job->vi = p.vi; job->vj = p.vj;
  job->tiles = (static_cast<uint32_t>(t0i) * static_cast<uint32_t>(g.TJ) + static_cast<uint32_t>(t0j)) * kTile;
  job->box = static_cast<uint32_t>(i0 - ci0) | (static_cast<uint32_t>(i1 - ci0) << 7) |
             (static_cast<uint32_t>(j0 - cj0) << 14) | (static_cast<uint32_t>(j1 - cj0) << 21);
  job->occ = occ;
  return true;
}

// Window of the warp: 96 rows x 96 columns of occupancy bits as sm[wcol * 96 + row]
// (wcol = column / 32).  Each staged tile is one coalesced 128 B line.
constexpr int kWinWords = 3 * 96;

// The three tiles of one tile row are 3 x 128 B contiguous in memory, so a row of the window is
// one 128-bit load by 24 lanes (lane l: tile l / 8, rows 4 (l % 8) .. +3) and one 128-bit store.
// Tile rows without an obstacle under the box are zero-filled instead of loaded.  Tiles beside
// the box may be loaded as well; no lattice sample can land there.  (The tile array is padded
// by two tiles so that the last tile row can be read three tiles wide.)
__device__ __forceinline__ void stage_window(const GridDev& g, const NarrowJob& job, uint32_t* sm, int lane) {
  const int b = lane >> 3, r4 = (lane & 7) * 4;
  // one 64-bit address for the first tile row; the next rows are TJ tiles (= TJ * 128 bytes) further
  const uint4* row = reinterpret_cast<const uint4*>(g.occ_tiles + job.tiles) + lane;
  const unsigned row_stride = static_cast<unsigned>(g.TJ) * (kTile / 4);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 24 && ((job.occ >> (a * 3)) & 7u)) w = __ldg(row + a * row_stride);
    if (lane < 24) *reinterpret_cast<uint4*>(&sm[b * 96 + a * 32 + r4]) = w;
  }
  __syncwarp();
}

// The same staging as asynchronous copies (LDGSTS, 16 bytes per lane and tile row, zero-filled where the tile row
// holds nothing): the caller commits the group and waits for it when it gets to the job, so the L2 round trip of
// job j + 1 .. j + depth runs under the lattice of job j.
__device__ __forceinline__ void stage_window_async(const GridDev& g, const NarrowJob& job, uint32_t* sm, int lane) {
  const int b = lane >> 3, r4 = (lane & 7) * 4;
  const uint4* row = reinterpret_cast<const uint4*>(g.occ_tiles + job.tiles) + lane;
  const unsigned row_stride = static_cast<unsigned>(g.TJ) * (kTile / 4);
  if (lane < 24) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const bool have = (job.occ >> (a * 3)) & 7u;
      const uint4* src = have ? row + a * row_stride : reinterpret_cast<const uint4*>(g.occ_tiles);   // (never read when !have)
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(&sm[b * 96 + a * 32 + r4]));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(have ? 16 : 0) : "memory");
    }
  }
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// exact "any occupied cell inside the lattice's bounding box" on the staged window (the light path's pre-test)
__device__ __forceinline__ bool window_box_any(const NarrowJob& job, const uint32_t* sm, int lane) {
  const int ri0 = job.box & 127, ri1 = (job.box >> 7) & 127;
  const int rj0 = (job.box >> 14) & 127, rj1 = (job.box >> 21) & 127;
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

// "Any occupied cell the footprint can reach" on the staged window, cell row by cell row (lane l: rows l, l + 32,
// l + 64): the job's bands give the column range of each row.  False proves the pose free; true sends it to the
// lattice.  (The bounding box alone -- what this test replaced -- lets through 2.6 times the footprint's area.)
__device__ __forceinline__ bool window_hull_any(const NarrowJob& job, const uint32_t* sm, int lane) {
  const int ri0 = job.box & 127, ri1 = (job.box >> 7) & 127;
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int r = lane + 32 * k;
    const float d = static_cast<float>(r) + job.d0;
    const float lo = fmaxf(fmaxf(fmaf(job.kT, d, job.tlo), fmaf(job.kS, d, job.slo)), 0.0f);
    const float hi = fminf(fminf(fmaf(job.kT, d, job.thi), fmaf(job.kS, d, job.shi)), 95.0f);
    if (r >= ri0 && r <= ri1 && lo <= hi) {
      const int a = static_cast<int>(lo), b = static_cast<int>(hi);
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        const int lw = max(a - 32 * w, 0), hw = min(b - 32 * w, 31);
        if (lw <= hw) any |= sm[96 * w + r] & (0xffffffffu >> (31 - hw)) & (0xffffffffu << lw);
      }
    }
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

// A cheap first look along the lattice's PERIMETER: 32 samples spread over each long side (b = 0 and b = NV-1) and
// 16 over each end (a = 0 and a = NU-1), three lookups per lane.  A solid obstacle wider than the car cannot overlap
// the footprint without crossing its border, so nearly every real collision is proven here at a tenth of the cost of
// the full lattice.  Every sample is a lattice sample, so a hit here is a hit of the full test; no hit proves
// nothing and the caller goes on to the full lattice (thin or small obstacles, and the free poses).
template <int NU, int NV>
__device__ __forceinline__ bool lattice_perimeter_any(const NarrowJob& job, const uint32_t* sm, int lane) {
  const int a = (lane * (NU - 1)) / 31;                    // 0 .. NU-1 over the 32 lanes
  const int eb = ((lane & 15) * (NV - 1)) / 15;            // 0 .. NV-1 over each half warp
  const int ea = (lane & 16) ? NU - 1 : 0;
  const int32_t si = job.oi + a * job.ui, sj = job.oj + a * job.uj;
  const int32_t ei = job.oi + ea * job.ui + eb * job.vi, ej = job.oj + ea * job.uj + eb * job.vj;
  const uint32_t f[3][2] = {{static_cast<uint32_t>(si), static_cast<uint32_t>(sj)},
                            {static_cast<uint32_t>(si + (NV - 1) * job.vi), static_cast<uint32_t>(sj + (NV - 1) * job.vj)},
                            {static_cast<uint32_t>(ei), static_cast<uint32_t>(ej)}};
  uint32_t hit = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t w = sm[(f[k][1] >> 29) * 96 + (f[k][0] >> kFracBits)];
    hit |= w >> ((f[k][1] >> kFracBits) & 31u);
  }
  return __any_sync(0xffffffffu, (hit & 1u) != 0);
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


// The narrow phase proper on a staged window (jobs of the broad phase): a cheap pass along the lattice's perimeter
// proves nearly every collision; what it does not prove goes through the hull test, and only then to the full
// lattice.  NU == 0: run-time lattice shape.
template <int NU, int NV, int S>
__device__ __forceinline__ bool window_collide_fixed(const Lattice& l, const NarrowJob& job, const uint32_t* sm, int lane) {
  if (NU > 0) {
    if (lattice_perimeter_any<(NU > 0 ? NU : 4), (NU > 0 ? NV : 4)>(job, sm, lane)) return true;
    if (!window_hull_any(job, sm, lane)) return false;
    return lattice_any_fixed<(NU > 0 ? NU : 1), (NU > 0 ? NV : 1), S>(job, sm, lane);
  }
  if (!window_hull_any(job, sm, lane)) return false;
  return lattice_any_generic(l, job, sm, lane);
}

// the light path's narrow phase on a staged window: box test, then the lattice (run-time dispatch on its shape)
__device__ __forceinline__ bool window_box_collide(const Lattice& l, const NarrowJob& job, const uint32_t* sm, int lane) {
  if (!window_box_any(job, sm, lane)) return false;
  if (l.nu == 48 && l.nv == 17) return lattice_any_fixed<48, 17, 2>(job, sm, lane);
  if (l.nu == 20 && l.nv == 8) return lattice_any_fixed<20, 8, 4>(job, sm, lane);
  return lattice_any_generic(l, job, sm, lane);
}

// Up to 32 poses, one per lane (`need`: this lane's pose survived make_job_light and `job` describes it): the warp
// runs the narrow phase for them in turn and returns true as soon as one collides.  All lanes return the same value.
__device__ __forceinline__ bool warp_collide_jobs(const GridDev& g, const Lattice& lat, const NarrowJob& job,
                                                  bool need, uint32_t* sm, int lane) {
  uint32_t todo = __ballot_sync(0xffffffffu, need);
  bool hit = false;
  while (todo && !hit) {
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
    hit = window_box_collide(lat, j, sm, lane);
    __syncwarp();
  }
  return hit;
}

}  // namespace pp
