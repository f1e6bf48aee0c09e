// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// Philox4x32-10 counter-based generator, restated from the published algorithm
// (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11):
// ten rounds of two 32x32->64 multiplies, key bumped by the Weyl constants between
// rounds.  The reference repository has no RNG in its planner; this is
// row S1 of the scope table, an extension.  Known-answer vectors from the paper's
// reference implementation are pinned in tests/test_oracle_kat.py::test_philox_known_answers.
#pragma once
#include <cstdint>

namespace oracle {

struct Philox4 {
  uint32_t v[4];
};

inline Philox4 philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2]) {
  constexpr uint32_t kMul0 = 0xD2511F53u, kMul1 = 0xCD9E8D57u;
  constexpr uint32_t kWeyl0 = 0x9E3779B9u, kWeyl1 = 0xBB67AE85u;
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = static_cast<uint64_t>(kMul0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(kMul1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += kWeyl0;
    k1 += kWeyl1;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// 24 random bits -> float in [0, 1), exactly representable
inline float u01_from_bits(uint32_t r) { return static_cast<float>(r >> 8) * 0x1.0p-24f; }

// stream tags (third counter word is the caller's stream / query id, fourth is the tag)
constexpr uint32_t kTagPose = 0x504F5345u;  // "POSE": pose batches (config 3)
constexpr uint32_t kTagRrt = 0x52525421u;   // "RRT!": sampling planner draws

}  // namespace oracle
