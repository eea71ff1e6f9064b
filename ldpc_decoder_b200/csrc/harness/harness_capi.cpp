// harness_capi.cpp -- C hooks on the CLI harness' data generator for the ctypes tests (CPU only, no CUDA).
#include "chacha8.h"
#include "half.h"
#include "test_data.h"

#include <cstdint>
#include <string>

using namespace ldpc_b200;

extern "C" {

void ldpc_b200_harness_random_ints(uint64_t seed, uint32_t n, uint32_t *out) {
  prng_chacha r(seed);
  for (uint32_t i = 0; i < n; ++i) out[i] = r.random_int();
}
void ldpc_b200_harness_units(uint64_t seed, uint32_t n, float *out) {
  prng_chacha r(seed);
  for (uint32_t i = 0; i < n; ++i) out[i] = r.unit();
}
void ldpc_b200_harness_gaussians(uint64_t seed, uint32_t n, float *out) {
  prng_chacha r(seed);
  for (uint32_t i = 0; i < n; ++i) out[i] = r.gaussian();
}
float ldpc_b200_harness_capacity(int kind, float noise) { return channel_capacity(kind, noise); }
uint16_t ldpc_b200_harness_float_to_half(float f) { return float_to_half_bits(f); }
float ldpc_b200_harness_half_to_float(uint16_t h) { return half_bits_to_float(h); }

int ldpc_b200_harness_create_data(const char *alist, int kind, float noise, uint32_t start_idx, uint32_t n_vec,
                                  uint32_t batch_idx, int fp16, void *noisy, uint32_t *ref_frames, uint32_t *syndromes) {
  try {
    ldpc_code code{std::string(alist)};
    create_data(code, kind, noise, start_idx, n_vec, batch_idx, fp16 != 0, noisy, ref_frames, syndromes);
    return 0;
  } catch (...) {
    return 1;
  }
}
}
