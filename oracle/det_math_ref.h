// TEST INFRASTRUCTURE -- CPU oracle.  Not part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/.
//
// det_math_ref.h: deterministic sin / cos / atan2 in explicit IEEE-754 binary64
// operations (+, -, *, /, fma, rint).  Every operation is individually rounded
// (compile with -ffp-contract=off), so the same sequence on the GPU (csrc/pp_math.cuh,
// written with __dmul_rn/__dadd_rn/__fma_rn) produces the SAME BITS.  glibc's and CUDA's
// libm differ in the last ulp, and the reference planner compares doubles produced by
// cos/sin with == (graph_node.hpp:24-30) and orders its frontier by costs that contain
// atan2 (LP:127), so bit-exact CPU<->GPU parity needs one shared definition of these three
// functions.  Accuracy against glibc is checked in tests/test_oracle_math.py (<= 2 ulp).
//
// This file is an independent restatement of the arithmetic in csrc/pp_math.cuh; the two
// are kept in lock-step by the parity tests, not by sharing source.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {
namespace detm {

// pi/2 split in three parts (33 + 33 + 53 bits), 2/pi
constexpr double kPio2Hi = 0x1.921fb54400000p+0;
constexpr double kPio2Mid = 0x1.0b4611a600000p-34;
constexpr double kPio2Lo = 0x1.3198a2e037073p-69;
constexpr double kTwoOverPi = 0x1.45f306dc9c883p-1;
// 2*pi split, 1/(2*pi)
constexpr double kTwoPiHi = 0x1.921fb54400000p+2;
constexpr double kTwoPiMid = 0x1.0b4611a600000p-32;
constexpr double kTwoPiLo = 0x1.3198a2e037073p-67;
constexpr double kInvTwoPi = 0x1.45f306dc9c883p-3;
// pi/4, pi/2, pi as hi + lo pairs for atan2 assembly
constexpr double kPio4H = 0x1.921fb54442d18p-1;
constexpr double kPio4L = 0x1.1a62633145c07p-55;
constexpr double kPio2H = 0x1.921fb54442d18p+0;
constexpr double kPio2L = 0x1.1a62633145c07p-54;
constexpr double kPiH = 0x1.921fb54442d18p+1;
constexpr double kPiL = 0x1.1a62633145c07p-53;

// minimax coefficients on [-pi/4, pi/4] (the classic fdlibm kernel polynomials)
constexpr double kS1 = -1.66666666666666324348e-01;
constexpr double kS2 = 8.33333333332248946124e-03;
constexpr double kS3 = -1.98412698298579493134e-04;
constexpr double kS4 = 2.75573137070700676789e-06;
constexpr double kS5 = -2.50507602534068634195e-08;
constexpr double kS6 = 1.58969099521155010221e-10;
constexpr double kC1 = 4.16666666666666019037e-02;
constexpr double kC2 = -1.38888888888741095749e-03;
constexpr double kC3 = 2.48015872894767294178e-05;
constexpr double kC4 = -2.75573143513906633035e-07;
constexpr double kC5 = 2.08757232129817482790e-09;
constexpr double kC6 = -1.13596475577881948265e-11;
// atan(z) = z - z^3 * P(z^2) on |z| <= 7/16
constexpr double kA[11] = {
    3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
    -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
    6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
    -3.65315727442169155270e-02, 1.62858201153657823623e-02};

inline void sincos(double x, double* s_out, double* c_out) {
  // quadrant: k = nearest integer to x * 2/pi (ties to even), r = x - k*pi/2
  const double k = std::rint(x * kTwoOverPi);
  double r = std::fma(-k, kPio2Hi, x);
  r = std::fma(-k, kPio2Mid, r);
  r = std::fma(-k, kPio2Lo, r);
  const double z = r * r;
  // sin(r) = r + r*z*(S1 + z*(S2 + ... ))
  double ps = kS6;
  ps = std::fma(ps, z, kS5);
  ps = std::fma(ps, z, kS4);
  ps = std::fma(ps, z, kS3);
  ps = std::fma(ps, z, kS2);
  ps = std::fma(ps, z, kS1);
  const double rz = r * z;
  const double sr = std::fma(rz, ps, r);
  // cos(r) = 1 + z*(-1/2 + z*(C1 + z*(C2 + ...)))
  double pc = kC6;
  pc = std::fma(pc, z, kC5);
  pc = std::fma(pc, z, kC4);
  pc = std::fma(pc, z, kC3);
  pc = std::fma(pc, z, kC2);
  pc = std::fma(pc, z, kC1);
  pc = std::fma(pc, z, -0.5);
  const double cr = std::fma(pc, z, 1.0);
  // k may be huge for absurd inputs; reduce it exactly before the int conversion
  const double k4 = k - 4.0 * std::rint(k * 0.25);  // in {-2,-1,0,1,2}
  const int q = static_cast<int>(k4) & 3;
  double s, c;
  switch (q) {
    case 0: s = sr; c = cr; break;
    case 1: s = cr; c = -sr; break;
    case 2: s = -sr; c = -cr; break;
    default: s = -cr; c = sr; break;
  }
  *s_out = s;
  *c_out = c;
}

inline double sin(double x) { double s, c; sincos(x, &s, &c); return s; }
inline double cos(double x) { double s, c; sincos(x, &s, &c); return c; }

// atan on [0, 1]
inline double atan_unit(double q) {
  double base_h = 0.0, base_l = 0.0, t = q;
  if (q > 0.4375) {  // atan(q) = pi/4 + atan((q-1)/(q+1))
    t = (q - 1.0) / (q + 1.0);
    base_h = kPio4H;
    base_l = kPio4L;
  }
  const double z = t * t;
  double p = kA[10];
  for (int i = 9; i >= 0; --i) p = std::fma(p, z, kA[i]);
  const double tz = t * z;
  const double a = std::fma(-tz, p, t);  // t - t*z*P(z)
  return base_h + (a + base_l);
}

inline double atan2(double y, double x) {
  if (x != x || y != y) return x + y;  // NaN in, NaN out
  const double ax = std::fabs(x), ay = std::fabs(y);
  double a;
  if (ax == 0.0 && ay == 0.0) {
    a = 0.0;
  } else if (ay <= ax) {
    a = atan_unit(ay / ax);
  } else {
    const double t = atan_unit(ax / ay);
    a = kPio2H - (t - kPio2L);
  }
  if (std::signbit(x)) a = kPiH - (a - kPiL);
  return std::signbit(y) ? -a : a;
}

// d wrapped to about [-pi, pi]: d - 2*pi*rint(d / 2pi)
inline double wrap_pi(double d) {
  const double k = std::rint(d * kInvTwoPi);
  double r = std::fma(-k, kTwoPiHi, d);
  r = std::fma(-k, kTwoPiMid, r);
  r = std::fma(-k, kTwoPiLo, r);
  return r;
}

}  // namespace detm
}  // namespace oracle
