// C++14 caller written against the drop-in class (what the reference's do_test does, main.cpp:315-441):
// construct from an alist file + channel, ask parallel_factor(), decode, count errors.
#include "../ldpc_decoder_b200/csrc/ldpc_decoder_gpu.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace ldpc_b200;

// a channel the decoder has no conversion kernel for (decoding_input_is_llr() == true): the class converts the channel
// values with llr() on the host, as the reference's prepare_vectors does (ldpc_decoder_gpu.cu:209-215)
class custom_channel : public noisy_channel {
  float m_factor;

public:
  explicit custom_channel(float sigma) : m_factor(2 * (1 / (sigma * sigma))) {} // as biawgn_channel::factor()
  float llr(const float val) const override { return m_factor * val; }
  channelType channel() const override { return groupGauss; }
};

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  try {
    ldpc_code code(argv[1]);
    const float sigma = argc > 2 ? float(atof(argv[2])) : 0.6f;
    biawgn_channel channel(sigma);
    ldpc_decoder_gpu_static_parameters sp;
    sp.m_max_log_parallel_factor_user = 4;
    ldpc_decoder_gpu dec(code, channel, sp);
    ldpc_decoder_gpu_dynamic_parameters dp;
    dp.m_num_iter_max = 60;
    const uint32_t n = dec.parallel_factor() * 3 + 5;
    const int64_t N = code.n_inputs(), n_tx = N - code.n_erased_inputs(), S = (code.n_outputs() + 31) >> 5, W = N >> 5;
    std::mt19937 rng(7);
    std::normal_distribution<float> gauss(0.f, sigma);
    std::vector<uint32_t> bits(size_t(n) * W), syn(size_t(n) * S), res(size_t(n) * W);
    std::vector<float> input(size_t(N) * n, 0.f);
    for (uint32_t v = 0; v < n; ++v) {
      for (int64_t w = 0; w < W; ++w) bits[size_t(v) * W + w] = rng();
      compute_syndrome(code, &bits[size_t(v) * W], &syn[size_t(v) * S]);
      for (int64_t i = 0; i < n_tx; ++i)
        input[size_t(v) + size_t(n) * i] = bool_to_llr((bits[size_t(v) * W + (i >> 5)] >> (i & 31)) & 1) + gauss(rng);
    }
    test_report report;
    if (dec.decoding_input_is_llr()) return 3;
    dec.decode(dp, n, input.data(), syn.data(), res.data(), report);
    uint64_t errors = 0;
    for (size_t i = 0; i < res.size(); ++i) errors += uint64_t(__builtin_popcount(res[i] ^ bits[i]));
    std::printf("frames %u errors %llu iter max/min/avg %u/%u/%.3f\n", n, (unsigned long long)errors, report.max_iter,
                report.min_iter, report.avg_iter);
    // extension: one queue of frames through two decoder objects (one per GPU on a multi-GPU box; here both on the
    // current device), a host thread each and a shared cursor inside the library
    {
      ldpc_decoder_gpu dec2(code, channel, sp);
      ldpc_decoder_gpu *both[2] = {&dec, &dec2};
      std::vector<uint32_t> res2(res.size(), 0u);
      test_report report2;
      ldpc_decoder_gpu::decode_multi(both, 2, dp, n, input.data(), syn.data(), res2.data(), report2);
      uint64_t errors2 = 0;
      for (size_t i = 0; i < res2.size(); ++i) errors2 += uint64_t(__builtin_popcount(res2[i] ^ bits[i]));
      std::printf("multi: frames %llu errors %llu iter max/min/avg %u/%u/%.3f\n",
                  (unsigned long long)report2.device_report.frames_decoded, (unsigned long long)errors2, report2.max_iter,
                  report2.min_iter, report2.avg_iter);
      if (errors2 != 0 || report2.device_report.frames_decoded != n) return 6;
    }
    // a custom channel: same LLRs as the AWGN channel, computed by channel.llr() on the host
    {
      custom_channel custom(sigma);
      ldpc_decoder_gpu dec3(code, custom, sp);
      if (!dec3.decoding_input_is_llr()) return 7;
      std::vector<uint32_t> res3(res.size(), 0u);
      test_report report3;
      dec3.decode(dp, n, input.data(), syn.data(), res3.data(), report3, 1);
      if (res3 != res || report3.max_iter != report.max_iter || report3.min_iter != report.min_iter) return 8;
      std::printf("custom channel: identical to the AWGN decode\n");
    }
    // malformed input must surface as the reference's exception type
    try {
      ldpc_code bad("/nonexistent.alist");
      return 4;
    } catch (const error &) {
    }
    return errors == 0 ? 0 : 1;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 5;
  }
}
