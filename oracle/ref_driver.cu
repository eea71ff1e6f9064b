// ref_driver.cu -- TEST INFRASTRUCTURE.  A thin C ABI around the UNMODIFIED reference CUDA backend
// (kunzjacq/ldpc_decoder: src/ldpc_decoder_gpu.cu, src/cuda/flood.cu, src/ldpc_code.cpp, ...), compiled
// from /root/reference where the sources lie (oracle/Makefile target `ref`) into oracle/_ref/.
// It lets the GPU parity tests run the reference's own decode() and its individual kernels on the same
// inputs as the B200 decoder, and generates the fixtures under tests/golden/.
//
// The reference's kernels and members are reached through its own headers; `private` is widened only to
// call prepare_vectors/transfer_vectors (h/ldpc_decoder_gpu_cuda.h:133-144) for kernel-level checks.
#define private public
#define protected public
#include "ldpc_decoder_gpu_cuda.h"
#undef private
#undef protected
#include "flood.cuh"

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

struct ref_handle {
  std::unique_ptr<ldpc_code> code;
  std::unique_ptr<noisy_channel> channel;
  std::unique_ptr<ldpc_decoder_gpu_cuda> dec;
};

static std::string g_err;

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
int ref_llr_size() { return (int)sizeof(llr_t); }

// channel_kind: 0 = AWGN (noise = sigma), 1 = BSC (noise = p)   [main.cpp:196-215]
void *ref_create(const char *alist_path, int channel_kind, float noise, uint32_t max_log2_frames) {
  try {
    std::unique_ptr<ref_handle> h(new ref_handle);
    h->code.reset(new ldpc_code(std::string(alist_path)));
    if (channel_kind == 1) h->channel.reset(new bsc_channel(noise));
    else h->channel.reset(new biawgn_channel(noise));
    ldpc_decoder_gpu_static_parameters sp;
    sp.m_max_log_parallel_factor_user = max_log2_frames;
    h->dec.reset(new ldpc_decoder_gpu_cuda(*h->code, *h->channel, sp));
    return h.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void ref_destroy(void *hv) { delete static_cast<ref_handle *>(hv); }

uint32_t ref_parallel_factor(void *hv) { return static_cast<ref_handle *>(hv)->dec->parallel_factor(); }

float ref_channel_param(void *hv) {
  ref_handle *h = static_cast<ref_handle *>(hv);
  if (h->channel->channel() == bsc) return dynamic_cast<bsc_channel *>(h->channel.get())->ref_llr();
  return dynamic_cast<biawgn_channel *>(h->channel.get())->factor();
}

// the reference's decode() (src/ldpc_decoder_gpu.cu:283-634); out4 = {max_iter, min_iter, avg_iter, iter_time}
int ref_decode(void *hv, uint32_t num_iter_max, uint32_t num_iter_check, uint32_t n_frames, void *input,
               const uint32_t *syndromes, uint32_t *results, float *out4) {
  ref_handle *h = static_cast<ref_handle *>(hv);
  try {
    ldpc_decoder_gpu_dynamic_parameters dp;
    dp.m_num_iter_max = num_iter_max;
    dp.m_num_iter_check_parity = num_iter_check;
    test_report rep;
    h->dec->decode(dp, n_frames, input, syndromes, results, rep, 0);
    if (out4) {
      out4[0] = (float)rep.max_iter;
      out4[1] = (float)rep.min_iter;
      out4[2] = rep.avg_iter;
      out4[3] = rep.iter_time_per_vector;
    }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

// ---- kernel-level steps on the reference's own buffers ------------------------------------------
// staged[j + cnt*i]: what prepare_vectors leaves in m_llrs; then transfer_vectors (:218-273)
int ref_refill(void *hv, uint32_t cnt, const void *staged, const uint32_t *syndromes) {
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  const int64_t n_reg = d->m_n_inputs - d->m_n_erased_inputs;
  memcpy(d->m_llrs, staged, sizeof(transfer_llr_t) * n_reg * cnt);
  d->transfer_vectors(cnt, d->m_llrs, syndromes);
  return (int)cudaDeviceSynchronize();
}

int ref_backward(void *hv) { // :347
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  flood_backward<<<d->m_tiles, d->m_local_threads>>>(d->m_syndrome_dev_buf, d->m_message_dev_buf,
                                                     d->m_out_bit_to_edge_dev_buf, d->m_syndrome_uint32_sz,
                                                     d->m_log2_parallel_factor, d->m_n_outputs,
                                                     d->m_log2_threads_per_vector);
  return (int)cudaDeviceSynchronize();
}

int ref_forward(void *hv, int final_bits) { // :353 / :362
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  if (final_bits)
    flood_forward_w_final_bits<<<d->m_tiles, d->m_local_threads>>>(
        d->m_message_dev_buf, d->m_initial_llrs_dev_buf, d->m_in_to_out_edge_dev_buf, d->m_in_bit_to_edge_dev_buf,
        d->m_final_bits_dev_buf, d->m_log2_parallel_factor, d->m_n_inputs, d->m_log2_global_threads);
  else
    flood_forward<<<d->m_tiles, d->m_local_threads>>>(d->m_message_dev_buf, d->m_initial_llrs_dev_buf,
                                                      d->m_in_to_out_edge_dev_buf, d->m_in_bit_to_edge_dev_buf,
                                                      d->m_log2_parallel_factor, d->m_n_inputs,
                                                      d->m_log2_global_threads);
  return (int)cudaDeviceSynchronize();
}

int ref_check(void *hv, uint8_t *violated) { // :367-375
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  cudaMemset(d->m_parities_violated_dev_buf, 0, d->m_parallel_factor);
  check_parity<<<d->m_tiles, d->m_local_threads>>>(d->m_syndrome_dev_buf, d->m_out_bit_to_edge_dev_buf,
                                                   d->m_out_edge_to_in_bit_dev_buf, d->m_final_bits_dev_buf,
                                                   d->m_parities_violated_dev_buf, d->m_syndrome_uint32_sz,
                                                   d->m_log2_parallel_factor, d->m_n_outputs,
                                                   d->m_log2_threads_per_vector);
  cudaMemcpy(violated, d->m_parities_violated_dev_buf, d->m_parallel_factor, cudaMemcpyDeviceToHost);
  return (int)cudaDeviceSynchronize();
}

int ref_deinterlace(void *hv, uint32_t *packed /* [P][N/32] */) { // :437-443
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  deinterlace_output<<<d->m_tiles, d->m_local_threads>>>(d->m_final_bits_dev_buf, d->m_final_bits_packed_dev_buf,
                                                         d->m_log2_parallel_factor, d->m_n_inputs,
                                                         d->m_log2_threads_per_vector);
  cudaMemcpy(packed, d->m_final_bits_packed_dev_buf, sizeof(uint32_t) * (d->m_n_inputs >> 5) * d->m_parallel_factor,
             cudaMemcpyDeviceToHost);
  return (int)cudaDeviceSynchronize();
}

// one slot's messages (check-major edge order) / priors, widened to float
static int read_col(const llr_t *dev, int64_t rows, uint32_t P, uint32_t slot, float *out) {
  std::vector<llr_t> all((size_t)rows * P);
  cudaError_t e = cudaMemcpy(all.data(), dev, sizeof(llr_t) * all.size(), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return (int)e;
  for (int64_t i = 0; i < rows; i++) {
#ifdef USE_FLOAT16_COMPUTE
    out[i] = __half2float(all[(size_t)i * P + slot]);
#else
    out[i] = all[(size_t)i * P + slot];
#endif
  }
  return 0;
}

int ref_read_messages(void *hv, uint32_t slot, float *out) {
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  return read_col(d->m_message_dev_buf, d->m_n_edges, d->m_parallel_factor, slot, out);
}

int ref_read_priors(void *hv, uint32_t slot, float *out) {
  ldpc_decoder_gpu_cuda *d = static_cast<ref_handle *>(hv)->dec.get();
  return read_col(d->m_initial_llrs_dev_buf, d->m_n_inputs, d->m_parallel_factor, slot, out);
}

// graph tables as the reference decoder built them (ldpc_decoder_gpu.cu:42-65)
void ref_graph(void *hv, int64_t *mne /* M,N,E,n_erased,maxdeg_in,maxdeg_out */, uint32_t *row_ptr, uint32_t *col_ptr,
               uint32_t *in_to_out, uint32_t *out_edge_to_var) {
  ref_handle *h = static_cast<ref_handle *>(hv);
  ldpc_decoder_gpu_cuda *d = h->dec.get();
  if (mne) {
    mne[0] = d->m_n_outputs;
    mne[1] = d->m_n_inputs;
    mne[2] = d->m_n_edges;
    mne[3] = d->m_n_erased_inputs;
    mne[4] = h->code->max_degree_in();
    mne[5] = h->code->max_degree_out();
  }
  if (row_ptr) memcpy(row_ptr, d->m_out_bit_to_edge, 4 * (d->m_n_outputs + 1));
  if (col_ptr) memcpy(col_ptr, d->m_in_bit_to_edge, 4 * (d->m_n_inputs + 1));
  if (in_to_out) memcpy(in_to_out, d->m_in_to_out_edge, 4 * d->m_n_edges);
  if (out_edge_to_var) memcpy(out_edge_to_var, d->m_out_edge_to_in_bit, 4 * d->m_n_edges);
}

} // extern "C"
