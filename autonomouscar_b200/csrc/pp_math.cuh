// pp_math.cuh -- deterministic binary64 sin / cos / atan2 for the device.
//
// The reference planner compares doubles that came out of cos()/sin() with ==
// (graph_node.hpp:24-30) and orders its frontier by costs containing atan2()
// (local_planner.cpp:127), and CUDA's libm is not bit-identical to glibc's.  These three
// functions are therefore defined once as explicit sequences of individually rounded
// IEEE operations (__dmul_rn / __dadd_rn / __fma_rn / __ddiv_rn never contract), which a
// CPU can replay exactly.  Accuracy vs glibc: <= 1 ulp (sin, cos), <= 3 ulp (atan2).
#pragma once
#include <cuda_runtime.h>

namespace pp {
namespace dm {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }

// Cody-Waite constants: pi/2 and 2*pi in three pieces (33 + 33 + 53 significant bits)
#define PP_PIO2_HI 0x1.921fb54400000p+0
#define PP_PIO2_MID 0x1.0b4611a600000p-34
#define PP_PIO2_LO 0x1.3198a2e037073p-69
#define PP_TWO_OVER_PI 0x1.45f306dc9c883p-1
#define PP_TWOPI_HI 0x1.921fb54400000p+2
#define PP_TWOPI_MID 0x1.0b4611a600000p-32
#define PP_TWOPI_LO 0x1.3198a2e037073p-67
#define PP_INV_TWOPI 0x1.45f306dc9c883p-3
#define PP_PIO4_H 0x1.921fb54442d18p-1
#define PP_PIO4_L 0x1.1a62633145c07p-55
#define PP_PIO2_H 0x1.921fb54442d18p+0
#define PP_PIO2_L 0x1.1a62633145c07p-54
#define PP_PI_H 0x1.921fb54442d18p+1
#define PP_PI_L 0x1.1a62633145c07p-53

__device__ __forceinline__ void sincos(double x, double* s_out, double* c_out) {
  const double k = rint(mul(x, PP_TWO_OVER_PI));
  double r = fma(-k, PP_PIO2_HI, x);
  r = fma(-k, PP_PIO2_MID, r);
  r = fma(-k, PP_PIO2_LO, r);
  const double z = mul(r, r);
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double sr = fma(mul(r, z), ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  pc = fma(pc, z, -0.5);
  const double cr = fma(pc, z, 1.0);
  const double k4 = sub(k, mul(4.0, rint(mul(k, 0.25))));
  const int q = __double2int_rz(k4) & 3;
  double s = sr, c = cr;
  if (q == 1) { s = cr; c = -sr; }
  else if (q == 2) { s = -sr; c = -cr; }
  else if (q == 3) { s = -cr; c = sr; }
  *s_out = s;
  *c_out = c;
}
__device__ __forceinline__ double sin_(double x) { double s, c; sincos(x, &s, &c); return s; }
__device__ __forceinline__ double cos_(double x) { double s, c; sincos(x, &s, &c); return c; }

__device__ __forceinline__ double atan_unit(double q) {
  double base_h = 0.0, base_l = 0.0, t = q;
  if (q > 0.4375) {
    t = div(sub(q, 1.0), add(q, 1.0));
    base_h = PP_PIO4_H;
    base_l = PP_PIO4_L;
  }
  const double z = mul(t, t);
  double p = 1.62858201153657823623e-02;
  p = fma(p, z, -3.65315727442169155270e-02);
  p = fma(p, z, 4.97687799461593236017e-02);
  p = fma(p, z, -5.83357013379057348645e-02);
  p = fma(p, z, 6.66107313738753120669e-02);
  p = fma(p, z, -7.69187620504482999495e-02);
  p = fma(p, z, 9.09088713343650656196e-02);
  p = fma(p, z, -1.11111104054623557880e-01);
  p = fma(p, z, 1.42857142725034663711e-01);
  p = fma(p, z, -1.99999999998764832476e-01);
  p = fma(p, z, 3.33333333333329318027e-01);
  const double a = fma(-mul(t, z), p, t);
  return add(base_h, add(a, base_l));
}

__device__ __forceinline__ bool signbit_(double v) { return (__double2hiint(v) >> 31) != 0; }

__device__ __forceinline__ double atan2_(double y, double x) {
  if (x != x || y != y) return add(x, y);
  const double ax = fabs(x), ay = fabs(y);
  double a;
  if (ax == 0.0 && ay == 0.0) {
    a = 0.0;
  } else if (ay <= ax) {
    a = atan_unit(div(ay, ax));
  } else {
    const double t = atan_unit(div(ax, ay));
    a = sub(PP_PIO2_H, sub(t, PP_PIO2_L));
  }
  if (signbit_(x)) a = sub(PP_PI_H, sub(a, PP_PI_L));
  return signbit_(y) ? -a : a;
}

__device__ __forceinline__ double wrap_pi(double d) {
  const double k = rint(mul(d, PP_INV_TWOPI));
  double r = fma(-k, PP_TWOPI_HI, d);
  r = fma(-k, PP_TWOPI_MID, r);
  r = fma(-k, PP_TWOPI_LO, r);
  return r;
}

// C++ double -> int truncation with the out-of-range rule of the build: anything not
// strictly inside +-2^30 means "too far away" (the reference's conversion is UB there)
__device__ __forceinline__ bool trunc_to_int(double v, int* out) {
  if (!(fabs(v) < 1073741824.0)) return false;
  *out = __double2int_rz(v);
  return true;
}

}  // namespace dm
}  // namespace pp
