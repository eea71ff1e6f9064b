// half.h -- IEEE binary16 <-> float on the host without CUDA headers (C++14); round to nearest even.
#pragma once
#include <cstdint>
#include <cstring>

namespace ldpc_b200 {

inline uint16_t float_to_half_bits(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return uint16_t(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0u)); // inf / nan
  if (x >= 0x477FF000u) return uint16_t(sign | 0x7C00u);                                    // overflows to inf
  if (x < 0x33000001u) return uint16_t(sign);                                               // rounds to zero
  if (x < 0x38800000u) { // subnormal half
    const int shift = 126 - int(x >> 23);        // 14..24
    uint32_t m = (x & 0x7FFFFFu) | 0x800000u;    // 24-bit significand
    const uint32_t rem_bits = uint32_t(shift);   // bits shifted out
    const uint32_t q = m >> rem_bits;
    const uint32_t rem = m & ((1u << rem_bits) - 1u), halfway = 1u << (rem_bits - 1);
    return uint16_t(sign | (q + ((rem > halfway || (rem == halfway && (q & 1u))) ? 1u : 0u)));
  }
  uint32_t h = ((x - 0x38000000u) >> 13);
  const uint32_t rem = x & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h; // may carry into the exponent: still correct
  return uint16_t(sign | h);
}

inline float half_bits_to_float(uint16_t h) {
  const uint32_t sign = uint32_t(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, x;
  if (e == 0) {
    if (m == 0) {
      x = sign;
    } else {
      int k = 0;
      while (!(m & 0x400u)) {
        m <<= 1;
        ++k;
      }
      x = sign | uint32_t(113 - k) << 23 | (m & 0x3FFu) << 13;
    }
  } else if (e == 31) {
    x = sign | 0x7F800000u | m << 13;
  } else {
    x = sign | (e + 112) << 23 | m << 13;
  }
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}

} // namespace ldpc_b200
