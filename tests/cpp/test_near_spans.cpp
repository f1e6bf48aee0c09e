// CPU check of the table behind the ORIENTED batch kernel's first test (csrc/pp_near_spans.h): for poses anywhere in a
// 4x4-cell block and any yaw, every 4x4 block in which a sample of the oracle's lattice (oracle/collision_ref.h,
// pose_to_fixed) lands must lie inside the span that the table gives for that block row and the yaw's bin -- with
// the pose's block taken from a cell index that may be one off in either axis, as the kernel's float arithmetic may be.
// usage: test_near_spans <res> <n_bins> <n_poses>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../autonomouscar_b200/csrc/pp_near_spans.h"
#include "../../oracle/collision_ref.h"

int main(int argc, char** argv) {
  const double res = argc > 1 ? atof(argv[1]) : 0.1;
  const int n_bins = argc > 2 ? atoi(argv[2]) : 16;
  const long n_poses = argc > 3 ? atol(argv[3]) : 20000;
  const oracle::Lattice l = oracle::make_lattice(4.7, 1.586, res);   // the launch file's car (full_sim.launch)
  if (!oracle::lattice_supported(l)) { printf("unsupported lattice\n"); return 2; }
  int E = 0;
  std::vector<int16_t> t;
  pp::near_span_table(l.su * l.hu, l.sv * l.hv, n_bins, &E, &t);
  const int rows = 2 * E + 1, W = 2048, H = 2048, halo = 128;
  std::mt19937_64 rng(2018);
  std::uniform_real_distribution<double> pos(-60.0, 60.0), ang(-12.0, 12.0);
  long bad = 0, checked = 0, widest = 0;
  for (long k = 0; k < n_poses; ++k) {
    const float x = static_cast<float>(pos(rng)), y = static_cast<float>(pos(rng));
    float yaw = static_cast<float>(ang(rng));
    if (k % 5 == 0) yaw = static_cast<float>((static_cast<int>(k / 5) % (8 * n_bins) - 4 * n_bins) * (M_PI / n_bins));  // bin borders
    // the kernel's bin: floor(yaw * n_bins / pi) mod n_bins, in float
    const int bin = static_cast<int>(std::floor(yaw * (static_cast<float>(n_bins) / 3.14159265f))) & (n_bins - 1);
    const oracle::PoseFixed p = oracle::pose_to_fixed(l, W, H, res, x, y, yaw);
    if (!p.valid) continue;
    const int ic = p.ai + oracle::kAnchorOffset, jc = p.aj + oracle::kAnchorOffset;   // floor of the centre
    for (int oi = -1; oi <= 1; ++oi)
      for (int oj = -1; oj <= 1; ++oj) {            // the kernel's cell index, possibly one off
        const int r4 = (ic + oi + halo) >> 2, c4 = (jc + oj + halo) >> 2;
        for (int a = 0; a < l.nu; ++a)
          for (int b = 0; b < l.nv; ++b) {
            const int32_t fi = p.oi + a * p.ui + b * p.vi, fj = p.oj + a * p.uj + b * p.vj;
            const int gi = p.ai + (fi >> oracle::kFracBits), gj = p.aj + (fj >> oracle::kFracBits);
            const int di = ((gi + halo) >> 2) - r4, dj = ((gj + halo) >> 2) - c4;
            ++checked;
            if (di < -E || di > E) { ++bad; continue; }
            const int first = t[(static_cast<size_t>(bin) * rows + di + E) * 2], last = t[(static_cast<size_t>(bin) * rows + di + E) * 2 + 1];
            if (dj < first || dj > last) ++bad;
          }
      }
  }
  for (int b = 0; b < n_bins; ++b)
    for (int r = 0; r < rows; ++r) {
      const int first = t[(static_cast<size_t>(b) * rows + r) * 2], last = t[(static_cast<size_t>(b) * rows + r) * 2 + 1];
      if (last >= first) widest = std::max<long>(widest, last - first + 1);
      if (last >= 60 || first <= -60) ++bad;        // the dilation kernel shifts within the neighbouring 64-bit words
    }
  printf("res %.3f bins %d lattice %dx%d E %d widest span %ld samples checked %ld outside %ld\n", res, n_bins, l.nu, l.nv, E,
         widest, checked, bad);
  return bad == 0 ? 0 : 1;
}
