// chacha8.h -- ChaCha8-based PRNG of the test harness (C++14, scalar).
// Produces the same streams as the reference's prng_chacha (h/prng_chacha.h, src/prng_chacha.cpp:18-67) and rng base
// (h/rng.h:7-71): key words 0/1 = seed low/high, 1536-byte refills, nonce = refill counter, block counter from 0,
// 8 rounds; unit() = (u32 + 0.5) * 2^-32 in float; gaussian() = Marsaglia polar method with a cached second value.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace ldpc_b200 {

class prng_chacha {
public:
  explicit prng_chacha(uint64_t seed) { reset_seed(seed); }

  void reset_seed(uint64_t seed) { // rng::reset_seed + prng_chacha::reset_seed_internal
    key_[0] = uint32_t(seed & 0xFFFFFFFFu);
    key_[1] = uint32_t(seed >> 32);
    for (int i = 2; i < 8; ++i) key_[i] = 0;
    nonce_ = 0;
    refill();
    g_val_exists_ = false;
  }

  uint32_t random_int() {
    if (idx_ == kWords) refill();
    return buf_[idx_++];
  }

  float unit() { // h/rng.h:38-42
    static const float normalizer = std::pow(2.f, -32.f);
    return (static_cast<float>(random_int()) + .5f) * normalizer;
  }

  float gaussian() { // h/rng.h:49-70
    float result;
    if (g_val_exists_) {
      result = g_val_next_;
    } else {
      float x, y, sqnorm;
      do {
        x = 2.f * unit() - 1.f;
        y = 2.f * unit() - 1.f;
        sqnorm = x * x + y * y;
      } while (sqnorm >= 1 || sqnorm == 0);
      const float modulus = std::sqrt((-2 * std::log(sqnorm)) / sqnorm);
      result = x * modulus;
      g_val_next_ = y * modulus;
    }
    g_val_exists_ = !g_val_exists_;
    return result;
  }

private:
  static constexpr unsigned kWords = 1536 / 4; // 24 blocks per refill (src/prng_chacha.cpp:31)

  static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
  static void quarter(uint32_t *x, int a, int b, int c, int d) {
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
    x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
    x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
  }

  void refill() { // chacha_avx(out, 1536, nonce, key, counter = 0, rounds = 8), src/chacha_stream.cpp:103-234
    for (uint32_t blk = 0; blk < kWords / 16; ++blk) {
      uint32_t in[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u,
                         key_[0], key_[1], key_[2], key_[3], key_[4], key_[5], key_[6], key_[7],
                         blk, 0u, uint32_t(nonce_ & 0xFFFFFFFFu), uint32_t(nonce_ >> 32)};
      uint32_t x[16];
      std::memcpy(x, in, sizeof(x));
      for (int r = 0; r < 4; ++r) { // 8 rounds = 4 double rounds
        quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
        quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
      }
      for (int i = 0; i < 16; ++i) buf_[blk * 16 + i] = x[i] + in[i];
    }
    idx_ = 0;
    ++nonce_;
  }

  uint32_t buf_[kWords];
  uint32_t key_[8];
  uint64_t nonce_ = 0;
  unsigned idx_ = 0;
  bool g_val_exists_ = false;
  float g_val_next_ = 0.f;
};

} // namespace ldpc_b200
