// ldpc_decoder_gpu.h -- header-compatible C++14 decoder class over the C ABI (include/ldpc_b200.h).
//
// Drop-in for the reference's ldpc_decoder_gpu / ldpc_decoder_gpu_cuda
// (h/ldpc_decoder_gpu.h:15-139, h/ldpc_decoder_gpu_cuda.h:11-145): same constructor arguments, same
// decode() signature and buffer contracts, same parallel_factor() / decoding_input_is_llr() /
// set_erased_variables().  Errors of the C ABI are re-thrown as `error` (h/common.h:61-73).
//
// Differences a caller can observe (see DESIGN.md "Deliberate deviations"):
//  * precision is a constructor-time choice (m_use_float16) instead of a compile-time macro; it defaults to
//    the reference's macro USE_FLOAT16_COMPUTE when that is defined;
//  * CUDA failures throw instead of calling exit() (cuda_manager.h:13-19).
#pragma once

#include "../../include/ldpc_b200.h"
#include "channel.h"
#include "common.h"
#include "ldpc_code.h"

#include <cstdint>
#include <string>

namespace ldpc_b200 {

// h/ldpc_decoder_gpu_common.h:7-24
class ldpc_decoder_gpu_static_parameters {
public:
  uint32_t m_max_log_parallel_factor_user; // log2 of the frames-in-flight requested (memory may lower it)
  int m_log2_local_threads;                // accepted for source compatibility; the kernels size themselves
  int m_log2_global_threads;               //  "
  bool m_use_float16;                      // extension: fp16 messages (reference: -DUSE_FLOAT16_COMPUTE)
  int m_device;                            // extension: CUDA device ordinal, -1 = current
  ldpc_decoder_gpu_static_parameters()
      : m_max_log_parallel_factor_user(5), m_log2_local_threads(9), m_log2_global_threads(25),
#ifdef USE_FLOAT16_COMPUTE
        m_use_float16(true),
#else
        m_use_float16(false),
#endif
        m_device(-1) {
  }
};

// h/ldpc_decoder_gpu_common.h:26-54
class ldpc_decoder_gpu_dynamic_parameters {
public:
  float m_infinity_threshold; // only read by the reference's OpenCL kernels; unused here
  uint32_t m_num_iter_max;
  uint32_t m_num_iter_check_parity;
  uint32_t m_num_vectors_per_run;
  uint32_t m_loading_factor;
  uint32_t m_target_errors;
  ldpc_decoder_gpu_dynamic_parameters()
      : m_infinity_threshold(10), m_num_iter_max(100), m_num_iter_check_parity(10), m_num_vectors_per_run(0),
        m_loading_factor(4), m_target_errors(0) {}
};

// the fields decode() writes (h/test_report.h:20-23); the harness' formatting lives with the harness
struct test_report {
  float avg_iter = 0;
  float iter_time_per_vector = 0;
  uint32_t min_iter = uint32_t(-1);
  uint32_t max_iter = 0;
  // extension: what the B200 path measured for this call
  ldpc_b200_report device_report = ldpc_b200_report();
};

class ldpc_decoder_gpu {
public:
  ldpc_decoder_gpu(const ldpc_code &p_code, const noisy_channel &p_channel,
                   const ldpc_decoder_gpu_static_parameters &p_params)
      : m_channel(p_channel) {
    ldpc_b200_code *code = nullptr;
    check(ldpc_b200_code_from_csr(uint32_t(p_code.n_outputs()), uint32_t(p_code.n_inputs()), p_code.row_ptr().data(),
                                  p_code.out_edge_to_var().data(), uint32_t(p_code.n_erased_inputs()), &code));
    ldpc_b200_config cfg = ldpc_b200_config();
    cfg.precision = p_params.m_use_float16 ? LDPC_B200_FP16 : LDPC_B200_FP32;
    cfg.max_log2_frames = p_params.m_max_log_parallel_factor_user;
    cfg.device = p_params.m_device;
    const ldpc_b200_status st = ldpc_b200_decoder_create(code, &cfg, &m_dec);
    ldpc_b200_code_destroy(code); // the decoder copied the tables (as the reference does, ldpc_decoder_gpu.cu:34-65)
    check(st);
    check(ldpc_b200_decoder_prepare_host_io(m_dec)); // this API is always fed from host arrays
    m_parallel_factor = ldpc_b200_decoder_frames_in_flight(m_dec);
  }
  ~ldpc_decoder_gpu() { ldpc_b200_decoder_destroy(m_dec); }

  ldpc_decoder_gpu() = delete;
  ldpc_decoder_gpu(const ldpc_decoder_gpu &) = delete;
  ldpc_decoder_gpu &operator=(const ldpc_decoder_gpu &) = delete;

  // p_input[v + num_vectors*i] = i-th channel value of vector v (llr_t: float, or IEEE half when fp16);
  // p_syndromes frame-major u32[ceil(M/32)] per vector; p_results frame-major u32[N/32] per vector, LSB first.
  void decode(const ldpc_decoder_gpu_dynamic_parameters &p_dyn_params, uint32_t p_num_vectors_to_process,
              void *p_input, const uint32_t *p_syndromes, uint32_t *p_results, test_report &report,
              uint32_t p_log = 0) {
    (void)p_log;
    ldpc_b200_decode_params prm = ldpc_b200_decode_params();
    switch (m_channel.channel()) { // transfer_vectors' dispatch, ldpc_decoder_gpu.cu:233-256
    case bsc: {
      const bsc_channel *cc = dynamic_cast<const bsc_channel *>(&m_channel);
      if (!cc) throw error("channel reports bsc but is not a bsc_channel");
      prm.channel = LDPC_B200_BSC;
      prm.channel_param = cc->ref_llr();
      break;
    }
    case awgn: {
      const biawgn_channel *cc = dynamic_cast<const biawgn_channel *>(&m_channel);
      if (!cc) throw error("channel reports awgn but is not a biawgn_channel");
      prm.channel = LDPC_B200_AWGN;
      prm.channel_param = cc->factor();
      break;
    }
    default: // decoding_input_is_llr(): the caller already converted with channel.llr()
      prm.channel = LDPC_B200_LLR;
      prm.channel_param = 0;
    }
    prm.num_iter_max = p_dyn_params.m_num_iter_max;
    prm.num_iter_check_parity = p_dyn_params.m_num_iter_check_parity;
    prm.input_layout = LDPC_B200_LAYOUT_INTERLEAVED;
    ldpc_b200_report rep;
    check(ldpc_b200_decode(m_dec, &prm, p_num_vectors_to_process, p_input, p_syndromes, p_results, &rep, nullptr));
    if (p_num_vectors_to_process == 0) return;
    report.max_iter = rep.max_iter;
    report.min_iter = rep.min_iter;
    report.avg_iter = rep.avg_iter;
    report.iter_time_per_vector = rep.iter_time_per_vector;
    report.device_report = rep;
  }

  bool decoding_input_is_llr() const { // h/ldpc_decoder_gpu.h:112-118
    const bool has_kernel_for_llr_conversion = (m_channel.channel() == bsc) || (m_channel.channel() == awgn);
    return !has_kernel_for_llr_conversion;
  }
  uint32_t parallel_factor() const { return m_parallel_factor; }
  void set_erased_variables(unsigned int p_n_erased_inputs) {
    ldpc_b200_decoder_set_erased_vars(m_dec, p_n_erased_inputs);
  }
  ldpc_b200_decoder *handle() { return m_dec; }

private:
  static void check(ldpc_b200_status st) {
    if (st != LDPC_B200_OK) throw error(std::string(ldpc_b200_last_error()));
  }
  ldpc_b200_decoder *m_dec = nullptr;
  uint32_t m_parallel_factor = 0;
  const noisy_channel &m_channel; // must outlive the decoder, as in the reference
};

// the CUDA-flavoured name the reference's main.cpp instantiates under CUDA_DECODER (main.cpp:315-319)
typedef ldpc_decoder_gpu ldpc_decoder_gpu_cuda;

} // namespace ldpc_b200
