// pp_philox.cuh -- Philox4x32-10 (Salmon et al., SC'11) for the device, in registers.
// Row S1 of the scope table: the reference planner itself has no RNG.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pp {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
#ifdef __CUDA_ARCH__
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t h0 = uint32_t(p0 >> 32), l0 = uint32_t(p0), h1 = uint32_t(p1 >> 32), l1 = uint32_t(p1);
#endif
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ float u01_from_bits(uint32_t r) {
  return static_cast<float>(r >> 8) * 0x1.0p-24f;  // exact: 24 bits fit the significand
}

constexpr uint32_t kTagPose = 0x504F5345u;  // pose batches
constexpr uint32_t kTagRrt = 0x52525421u;   // sampling-planner draws
#define PP_PI_F 3.14159274101257324219f
#define PP_TWOPI_F 6.28318548202514648438f

}  // namespace pp
