// ldpc_b200_dropin.hpp -- the reference's decoder class re-implemented over the C ABI (ldpc_b200.h), C++14.
//
// One class template carries the whole public surface of the reference's ldpc_decoder_gpu / ldpc_decoder_gpu_cuda
// (h/ldpc_decoder_gpu.h:77-128, h/ldpc_decoder_gpu_cuda.h:84-132): same constructor arguments, decode() signature and
// buffer contracts, parallel_factor(), decoding_input_is_llr(), set_erased_variables().  It is generic over the types
// of the caller's tree (code graph, channels, parameter structs, report, exception), so that
//   * include/ldpc_decoder_gpu.h instantiates it on the GLOBAL-namespace types of the reference's own headers: the
//     reference's src/main.cpp compiles against it unchanged (only the include path changes), and
//   * ldpc_decoder_b200/csrc/ldpc_decoder_gpu.h instantiates it on this repository's C++14 mirror types
//     (namespace ldpc_b200), for callers that do not have the reference tree.
// Errors of the C ABI are re-thrown as the tree's `error` type (h/common.h:61-73).
#ifndef LDPC_B200_DROPIN_HPP
#define LDPC_B200_DROPIN_HPP

#include "ldpc_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace ldpc_b200_dropin {

// ---- IEEE half <-> float for the "input already holds LLRs" path of an fp16 build (C++14 has no half type) ----
inline float half_bits_to_float(uint16_t h) {
  const uint32_t sign = uint32_t(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { // subnormal: normalise
      int e = -1;
      do {
        ++e;
        man <<= 1;
      } while (!(man & 0x400u));
      bits = sign | uint32_t(127 - 15 - e) << 23 | (man & 0x3FFu) << 13;
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | man << 13;
  } else {
    bits = sign | (exp + 112) << 23 | man << 13;
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}
inline uint16_t float_to_half_bits(float f) { // round to nearest even
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint16_t sign = uint16_t((x >> 16) & 0x8000u);
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return uint16_t(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0u));
  if (x >= 0x477FF000u) return uint16_t(sign | 0x7C00u); // rounds to infinity
  if (x < 0x33000001u) return sign;                       // rounds to zero
  int32_t e = int32_t(x >> 23) - 127;
  uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
  uint32_t shift = e < -14 ? uint32_t(13 + (-14 - e)) : 13u;
  uint32_t half_man = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_man & 1u))) ++half_man;
  const uint32_t base = e < -14 ? 0u : uint32_t(e + 15 - 1) << 10; // half_man carries the implicit bit
  return uint16_t(sign | (base + half_man));
}

// precision of the build: a tree whose static-parameter class has an m_use_float16 member chooses at run time,
// otherwise the reference's compile-time macro decides (CMakeLists.txt:15,105-107)
template <class P> inline auto use_float16(const P &p, int) -> decltype(bool(p.m_use_float16)) { return p.m_use_float16; }
template <class P> inline bool use_float16(const P &, long) {
#ifdef USE_FLOAT16_COMPUTE
  return true;
#else
  return false;
#endif
}
template <class P> inline auto device_of(const P &p, int) -> decltype(int(p.m_device)) { return p.m_device; }
template <class P> inline int device_of(const P &, long) { return -1; }
// extension field of this repository's test_report mirror (the reference's has none)
template <class R> inline auto store_device_report(R &r, const ldpc_b200_report &rep, int) -> decltype((void)(r.device_report = rep)) {
  r.device_report = rep;
}
template <class R> inline void store_device_report(R &, const ldpc_b200_report &, long) {}

// Types bundle: { code_t, channel_t, bsc_t, awgn_t, static_params_t, dynamic_params_t, report_t, error_t,
//                 static const int bsc_id, awgn_id }  (ids = the tree's channelType enumerators)
template <class T> class decoder {
public:
  typedef typename T::code_t code_t;
  typedef typename T::channel_t channel_t;
  typedef typename T::static_params_t static_params_t;
  typedef typename T::dynamic_params_t dynamic_params_t;
  typedef typename T::report_t report_t;
  typedef typename T::error_t error_t;

  // ldpc_decoder_gpu.cu:20-157: validate the graph, size the frames in flight from device memory, allocate, upload
  decoder(const code_t &p_code, const channel_t &p_channel, const static_params_t &p_params) : m_channel(p_channel) {
    // a caller compiled against another version of ldpc_b200.h than the library would corrupt its own stack
    uint32_t abi[4] = {0, 0, 0, 0};
    ldpc_b200_abi_sizes(abi);
    if (abi[1] != sizeof(ldpc_b200_config) || abi[2] != sizeof(ldpc_b200_decode_params) || abi[3] != sizeof(ldpc_b200_report))
      throw error_t(std::string("ldpc_b200.h and libldpc_b200.so disagree on the struct layouts: rebuild the caller"));
    const uint32_t M = uint32_t(p_code.n_outputs()), N = uint32_t(p_code.n_inputs()), E = p_code.n_edges();
    std::vector<uint32_t> row_ptr(size_t(M) + 1), col_idx(E);
    for (uint32_t c = 0; c < M; ++c) row_ptr[c] = p_code.out_bit_to_edge(c);
    row_ptr[M] = E;
    // variable of a check-major edge (the reference's out_edge_to_in_bit, ldpc_decoder_gpu.cu:60-65)
    for (uint32_t e = 0; e < E; ++e) col_idx[e] = p_code.in_edge_to_bit(p_code.edge_out_to_in(e));
    m_fp16 = use_float16(p_params, 0);
    ldpc_b200_code *code = nullptr;
    check(ldpc_b200_code_from_csr(M, N, row_ptr.data(), col_idx.data(), uint32_t(p_code.n_erased_inputs()), &code));
    ldpc_b200_config cfg = ldpc_b200_config();
    cfg.precision = m_fp16 ? LDPC_B200_FP16 : LDPC_B200_FP32;
    cfg.max_log2_frames = p_params.m_max_log_parallel_factor_user;
    cfg.device = device_of(p_params, 0);
    ldpc_b200_status st = ldpc_b200_decoder_create(code, &cfg, &m_dec);
    ldpc_b200_code_destroy(code); // the decoder copied the tables (as the reference does, ldpc_decoder_gpu.cu:34-65)
    if (st == LDPC_B200_OK) st = ldpc_b200_decoder_prepare_host_io(m_dec); // this API is always fed from host arrays
    if (st != LDPC_B200_OK) { // the destructor does not run for a constructor that throws
      const std::string msg = ldpc_b200_last_error();
      ldpc_b200_decoder_destroy(m_dec);
      m_dec = nullptr;
      throw error_t(msg);
    }
    m_parallel_factor = ldpc_b200_decoder_frames_in_flight(m_dec);
    m_n_inputs = N;
    m_n_erased = uint32_t(p_code.n_erased_inputs());
  }
  ~decoder() { ldpc_b200_decoder_destroy(m_dec); }
  decoder() = delete;
  decoder(const decoder &) = delete;
  decoder &operator=(const decoder &) = delete;

  // p_input[v + num_vectors*i] = i-th channel value of vector v (transfer_llr_t: float, or IEEE half in an fp16
  // build); p_syndromes frame-major u32[ceil(M/32)] per vector; p_results frame-major u32[N/32] per vector, LSB
  // first (h/ldpc_decoder_gpu.h:88-110).  p_log: 1 = total time, 2 = + time of every step of the schedule, summed
  // over the call (the reference prints them per refill with a device synchronisation in between,
  // ldpc_decoder_gpu.cu:528-601), 3 = + one line per frame with its iteration count (:405-412).
  void decode(const dynamic_params_t &p_dyn_params, uint32_t p_num_vectors_to_process, void *p_input,
              const uint32_t *p_syndromes, uint32_t *p_results, report_t &report, uint32_t p_log = 0) {
    if (p_num_vectors_to_process == 0) return; // ldpc_decoder_gpu.cu:293-294
    decode_on(&m_dec, 1, p_dyn_params, p_num_vectors_to_process, p_input, p_syndromes, p_results, report, p_log);
  }

  // EXTENSION (the reference drives device 0 only, h/cuda_manager.h:52): the same call on several
  // decoders of this class, one per GPU (static parameter m_device), constructed from the same code and channel --
  // one queue of frames, one host thread per device, a shared atomic cursor (ldpc_b200_decode_multi); the arrays and
  // the report are those of decode().
  static void decode_multi(decoder *const *p_decoders, uint32_t p_num_decoders, const dynamic_params_t &p_dyn_params,
                           uint32_t p_num_vectors_to_process, void *p_input, const uint32_t *p_syndromes,
                           uint32_t *p_results, report_t &report, uint32_t p_log = 0) {
    if (!p_decoders || !p_num_decoders) throw error_t(std::string("decode_multi needs at least one decoder"));
    if (p_num_vectors_to_process == 0) return;
    std::vector<ldpc_b200_decoder *> handles(p_num_decoders);
    for (uint32_t i = 0; i < p_num_decoders; ++i) handles[i] = p_decoders[i]->m_dec;
    p_decoders[0]->decode_on(handles.data(), p_num_decoders, p_dyn_params, p_num_vectors_to_process, p_input, p_syndromes,
                             p_results, report, p_log);
  }

private:
  void decode_on(ldpc_b200_decoder *const *p_handles, uint32_t p_num_handles, const dynamic_params_t &p_dyn_params,
                 uint32_t p_num_vectors_to_process, void *p_input, const uint32_t *p_syndromes, uint32_t *p_results,
                 report_t &report, uint32_t p_log) {
    ldpc_b200_decode_params prm = ldpc_b200_decode_params();
    std::vector<uint8_t> converted; // host copy holding LLRs, only for channels without a device conversion
    const int ch = int(m_channel.channel());
    if (ch == T::bsc_id) { // transfer_vectors' dispatch, ldpc_decoder_gpu.cu:233-256
      const typename T::bsc_t *cc = dynamic_cast<const typename T::bsc_t *>(&m_channel);
      if (!cc) throw error_t(std::string("channel reports bsc but is not a bsc_channel"));
      prm.channel = LDPC_B200_BSC;
      prm.channel_param = float(cc->ref_llr());
    } else if (ch == T::awgn_id) {
      const typename T::awgn_t *cc = dynamic_cast<const typename T::awgn_t *>(&m_channel);
      if (!cc) throw error_t(std::string("channel reports awgn but is not a biawgn_channel"));
      prm.channel = LDPC_B200_AWGN;
      prm.channel_param = float(cc->factor());
    } else {
      // decoding_input_is_llr(): no device kernel for this channel, so the channel values are converted on the host
      // with channel.llr(), as prepare_vectors does (ldpc_decoder_gpu.cu:209-215), into a copy of the input
      prm.channel = LDPC_B200_LLR;
      prm.channel_param = 0;
      const size_t n_tx = size_t(m_n_inputs - m_n_erased), count = n_tx * p_num_vectors_to_process;
      converted.resize(count * (m_fp16 ? 2 : 4));
      const unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < n_thr; ++t)
        pool.emplace_back([&, t] {
          const size_t a = count * t / n_thr, b = count * (t + 1) / n_thr;
          if (m_fp16) {
            const uint16_t *src = static_cast<const uint16_t *>(p_input);
            uint16_t *dst = reinterpret_cast<uint16_t *>(converted.data());
            for (size_t i = a; i < b; ++i) dst[i] = float_to_half_bits(float(m_channel.llr(half_bits_to_float(src[i]))));
          } else {
            const float *src = static_cast<const float *>(p_input);
            float *dst = reinterpret_cast<float *>(converted.data());
            for (size_t i = a; i < b; ++i) dst[i] = float(m_channel.llr(src[i]));
          }
        });
      for (std::thread &t : pool) t.join();
      p_input = converted.data();
    }
    prm.num_iter_max = p_dyn_params.m_num_iter_max;
    prm.num_iter_check_parity = p_dyn_params.m_num_iter_check_parity;
    prm.input_layout = LDPC_B200_LAYOUT_INTERLEAVED;
    if (p_log >= 2) prm.flags |= LDPC_B200_FLAG_TIME_KERNELS;
    std::vector<uint32_t> iters(p_log >= 3 ? p_num_vectors_to_process : 0);
    ldpc_b200_report rep;
    if (p_num_handles == 1)
      check(ldpc_b200_decode(p_handles[0], &prm, p_num_vectors_to_process, p_input, p_syndromes, p_results, &rep,
                             p_log >= 3 ? iters.data() : nullptr));
    else
      check(ldpc_b200_decode_multi(p_handles, p_num_handles, &prm, p_num_vectors_to_process, p_input, p_syndromes,
                                   p_results, &rep, nullptr, p_log >= 3 ? iters.data() : nullptr, 0));
    report.max_iter = rep.max_iter;
    report.min_iter = rep.min_iter;
    report.avg_iter = rep.avg_iter;
    report.iter_time_per_vector = rep.iter_time_per_vector;
    store_device_report(report, rep, 0);
    if (p_log >= 1) {
      std::ios fmt(nullptr);
      fmt.copyfmt(std::cout);
      std::cout << std::fixed << std::setprecision(3);
      if (p_log >= 3)
        for (uint32_t j = 0; j < p_num_vectors_to_process; ++j)
          std::cout << "   real idx = " << j << "; iterations: " << iters[j] << std::endl;
      if (p_log >= 2) {
        // the steps of the reference's schedule (ldpc_decoder_gpu.cu:346-609), CUDA-event time summed over the call
        std::cout << "time: " << rep.cn_s + rep.cn_refill_s << std::endl << "check node passes (flood_backward)" << std::endl;
        std::cout << "time: " << rep.vn_s << std::endl << "variable node passes (flood_forward)" << std::endl;
        std::cout << "time: " << rep.check_s << std::endl << "parity checks" << std::endl;
        std::cout << "time: " << rep.retire_s << std::endl << "Steps 2-4 (select, pack and retrieve finished vectors)" << std::endl;
        std::cout << "time: " << rep.refill_s << std::endl << "Steps 5-7 (load new vectors)" << std::endl;
        std::cout << "global iterations: " << rep.global_iters << "; parity checks: " << rep.n_checks_run
                  << "; vectors loaded after the first batch: " << rep.frames_refilled << std::endl;
      }
      std::cout << "decoder: time = " << rep.elapsed_s << "; final transfer done" << std::endl; // :629-631
      std::cout.copyfmt(fmt);
    }
  }

public:
  bool decoding_input_is_llr() const { // h/ldpc_decoder_gpu.h:112-118
    const int ch = int(m_channel.channel());
    const bool has_kernel_for_llr_conversion = (ch == T::bsc_id) || (ch == T::awgn_id);
    return !has_kernel_for_llr_conversion;
  }
  uint32_t parallel_factor() const { return m_parallel_factor; }
  void set_erased_variables(unsigned int p_n_erased_inputs) {
    ldpc_b200_decoder_set_erased_vars(m_dec, p_n_erased_inputs);
    m_n_erased = p_n_erased_inputs;
  }
  ldpc_b200_decoder *handle() { return m_dec; }

private:
  static void check(ldpc_b200_status st) {
    if (st != LDPC_B200_OK) throw error_t(std::string(ldpc_b200_last_error()));
  }
  ldpc_b200_decoder *m_dec = nullptr;
  uint32_t m_parallel_factor = 0, m_n_inputs = 0, m_n_erased = 0;
  bool m_fp16 = false;
  const channel_t &m_channel; // must outlive the decoder, as in the reference
};

} // namespace ldpc_b200_dropin

#endif // LDPC_B200_DROPIN_HPP
