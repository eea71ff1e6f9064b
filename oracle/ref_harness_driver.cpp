// ref_harness_driver.cpp -- TEST INFRASTRUCTURE.  C ABI around the UNMODIFIED reference test-data generator pieces
// (h/rng.h, src/prng_chacha.cpp, src/chacha_stream.cpp, src/channel.cpp, h/bool_vec.h, src/ldpc_code.cpp), compiled
// from /root/reference by oracle/Makefile target `refharness` into oracle/_ref/libldpc_ref_harness.so (CPU only).
// Used by tests/golden/make_golden_harness.py to pin the B200 CLI harness' data generation.
// create_data() itself lives in the reference's main.cpp, which cannot be compiled here (it includes the OpenCL
// headers), so the calls below reproduce its body statement by statement (src/main.cpp:450-538) on the reference's
// own classes.
#include "bool_vec.h"
#include "channel.h"
#include "ldpc_code.h"
#include "prng_chacha.h"

#include <cstdint>
#include <cstring>
#include <memory>

extern "C" {

void refh_random_ints(uint64_t seed, uint32_t n, uint32_t *out) {
  prng_chacha r(seed);
  for (uint32_t i = 0; i < n; i++) out[i] = r.random_int();
}

void refh_units(uint64_t seed, uint32_t n, float *out) {
  prng_chacha r(seed);
  for (uint32_t i = 0; i < n; i++) out[i] = r.unit();
}

void refh_gaussians(uint64_t seed, uint32_t n, float *out) {
  prng_chacha r(seed);
  for (uint32_t i = 0; i < n; i++) out[i] = r.gaussian();
}

float refh_capacity(int channel_kind, float noise) {
  if (channel_kind == 0) return bsc_channel(noise).capacity();
  return biawgn_channel(noise).capacity();
}

// main.cpp:450-538 (non-EXTRA_CHANNELS path), fp32 build
int refh_create_data(const char *alist, int channel_kind, float noise, uint32_t start_idx, uint32_t n_vec,
                     uint32_t batch_idx, float *noisy, uint32_t *ref_frames, uint32_t *syndromes) {
  try {
    ldpc_code code{std::string(alist)};
    std::unique_ptr<noisy_channel> ch;
    if (channel_kind == 0) ch.reset(new bsc_channel(noise));
    else ch.reset(new biawgn_channel(noise));
    const int64_t vec_uint32_sz = (code.n_inputs() + 0x1F) >> 5;
    const int64_t syndrome_uint32_sz = (n_effective_outputs(code) + 0x1F) >> 5;
    bool_vec ref_vec(n_vec, code.n_inputs());
    bool_vec syn_batch(n_vec, syndrome_uint32_sz << 5);
    const int64_t vec_sz = code.n_inputs();
    const int64_t transmitted_sz = code.n_inputs() - code.n_erased_inputs();
    const uint32_t num_words = ref_vec.num_words_per_bit();
    const uint32_t num_vec_rounded = num_words * bool_t_bit_size;
    const uint64_t vec_start_idx = start_idx + batch_idx * n_vec;
    prng_chacha r(0);
    for (uint32_t v_group = 0; v_group < num_words; v_group++) {
      prng_chacha rr(vec_start_idx + v_group * bool_t_bit_size);
      for (int64_t i = 0; i < vec_sz; i++) ref_vec.word_ref(i * num_words + v_group) = rr.random_int();
    }
    for (uint32_t v = 0; v < n_vec; v++) {
      r.reset_seed((vec_start_idx + v) | (1uLL << 32));
      int64_t i;
      for (i = 0; i < transmitted_sz; i++) {
        transfer_llr_t val = bool_to_llr(ref_vec[v + i * num_vec_rounded]);
        noisy[v + (size_t)n_vec * i] = ch->add_noise(r, val);
      }
      for (; i < vec_sz; i++) noisy[v + (size_t)n_vec * i] = 0;
    }
    // deinterlace (main.cpp:273-299) without the AVX2 transpose: bit i of word igroup of frame v = variable 32*igroup+i
    for (uint32_t v = 0; v < n_vec; v++)
      for (int64_t w = 0; w < vec_uint32_sz; w++) {
        uint32_t x = 0;
        for (int i = 0; i < 32; i++) x |= (uint32_t)ref_vec.bit(v, (w << 5) + i) << i;
        ref_frames[w + (size_t)v * vec_uint32_sz] = x;
      }
    compute_syndrome(code, ref_vec, syn_batch);
    for (uint32_t v = 0; v < n_vec; v++)
      for (int64_t w = 0; w < syndrome_uint32_sz; w++) {
        uint32_t x = 0;
        for (int i = 0; i < 32; i++) x |= (uint32_t)syn_batch.bit(v, (w << 5) + i) << i;
        syndromes[w + (size_t)v * syndrome_uint32_sz] = x;
      }
    return 0;
  } catch (...) {
    return 1;
  }
}
}
