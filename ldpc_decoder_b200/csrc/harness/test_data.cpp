#include "test_data.h"

#include "chacha8.h"
#include "half.h"

#include <cmath>

namespace ldpc_b200 {

namespace {

float bsc_capacity(float p) { return 1 + p * (std::log2(p)) + (1 - p) * (std::log2(1 - p)); } // channel.cpp:70-73

float log_cosh(float x, float range) { // channel.cpp:76-81
  const float abs_x = std::abs(x);
  if (abs_x > range) return abs_x - std::log(2.f);
  return std::log(std::cosh(x));
}

float biawgn_capacity(float s, float step, float range) { // channel.cpp:83-101
  float c = 0.;
  if (s < static_cast<float>(0.001)) {
    c = 1.;
  } else {
    const float inv_s = 1 / s;
    const float sq_inv_s = inv_s * inv_s;
    const float norm_factor = step / (std::log(2.f) * std::sqrt(2. * M_PI));
    for (float x = -range; x < range; x += step) c += std::exp(-x * x / 2) * (sq_inv_s - log_cosh(x * inv_s + sq_inv_s, range));
    c *= norm_factor;
  }
  return c;
}

} // namespace

float channel_capacity(int kind, float noise) {
  return kind == 0 ? bsc_capacity(noise) : biawgn_capacity(noise, 0.05f, 16.f);
}

void create_data(const ldpc_code &code, int kind, float noise, uint32_t start_idx, uint32_t n_vec, uint32_t batch_idx,
                 bool fp16, void *noisy_v, uint32_t *ref_frames, uint32_t *syndromes) {
  const int64_t N = code.n_inputs(), n_tx = N - code.n_erased_inputs();
  const int64_t W = (N + 31) >> 5, S = ((code.n_outputs() - code.n_erased_outputs()) + 31) >> 5;
  const uint32_t n_groups = (n_vec + 31) / 32;
  const uint64_t first = uint64_t(start_idx) + uint64_t(batch_idx) * n_vec; // main.cpp:470
  // one PRNG word per (variable, group of 32 frames): bit b of the word belongs to frame 32*group + b (:476-487)
  std::vector<uint32_t> sliced(size_t(N) * n_groups);
  for (uint32_t grp = 0; grp < n_groups; ++grp) {
    prng_chacha r(first + uint64_t(grp) * 32);
    for (int64_t i = 0; i < N; ++i) sliced[size_t(i) * n_groups + grp] = r.random_int();
  }
  float *noisy_f = static_cast<float *>(noisy_v);
  uint16_t *noisy_h = static_cast<uint16_t *>(noisy_v);
  prng_chacha r(0);
  for (uint32_t v = 0; v < n_vec; ++v) {
    r.reset_seed((first + v) | (1ull << 32)); // :520-522
    uint32_t *frame = ref_frames + size_t(v) * W;
    for (int64_t w = 0; w < W; ++w) frame[w] = 0;
    for (int64_t i = 0; i < N; ++i) {
      const uint32_t bit = (sliced[size_t(i) * n_groups + (v >> 5)] >> (v & 31)) & 1u;
      frame[i >> 5] |= bit << (i & 31);
      float out = 0.f; // punctured: channel output 0 (:530)
      if (i < n_tx) {
        const float val = bit ? 1.f : -1.f; // bool_to_llr, common.h:56-59
        if (kind == 0) { // bsc_channel::add_noise, channel.cpp:26-30
          out = (r.unit() < noise) ? -val : val;
        } else if (!fp16) { // biawgn_channel::add_noise, channel.cpp:65-68
          out = val + r.gaussian() * noise;
        } else { // fp16 build: gaussian() returns float16_t before the multiply
          out = val + half_bits_to_float(float_to_half_bits(r.gaussian())) * noise;
        }
      }
      if (fp16) noisy_h[size_t(v) + size_t(n_vec) * i] = float_to_half_bits(out);
      else noisy_f[size_t(v) + size_t(n_vec) * i] = out;
    }
    compute_syndrome(code, frame, syndromes + size_t(v) * S);
  }
}

} // namespace ldpc_b200
