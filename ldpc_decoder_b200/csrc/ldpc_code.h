// ldpc_code.h -- host-side code graph model (C++14), the B200 decoder's equivalent of the
// reference's ldpc_code (h/ldpc_code.h:10-62, src/ldpc_code.cpp:10-286).
//
// Builds, from an alist file, the two edge-ordered layouts the kernels consume:
//   CSR, check-major   : out_bit_to_edge[M+1] + out_edge_to_in_bit[E]   ("out" = check side)
//   CSC, variable-major: in_bit_to_edge[N+1]  + edge_in_to_out[E]       ("in"  = variable side)
// Messages live in HBM in check-major edge order, so edge_in_to_out is the gather list of the
// variable-node pass.  In-edges of a variable are ordered by increasing check index, exactly as
// the reference assigns them (ldpc_code.cpp:139-151), which fixes the summation order.
#pragma once

#include "common.h"

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ldpc_b200 {

class ldpc_code {
public:
  // p_is_filename == false parses p_alist as the file's text (ldpc_code.cpp:31-37)
  explicit ldpc_code(const std::string &p_alist, bool p_is_filename = true);
  // CSR constructor (harness-generated codes; no reference equivalent)
  ldpc_code(uint32_t n_checks, uint32_t n_vars, const uint32_t *row_ptr, const uint32_t *col_idx,
            uint32_t n_erased_vars);
  ldpc_code(const ldpc_code &) = delete;
  ldpc_code &operator=(const ldpc_code &) = delete;

  void set_n_erased_in_bits(int32_t p_n_erased_bits) { m_n_erased_variables = p_n_erased_bits; }
  int64_t n_inputs() const { return m_n_inputs; }
  int64_t n_outputs() const { return m_n_outputs; }
  int64_t n_erased_inputs() const { return m_n_erased_variables; }
  int64_t n_erased_outputs() const { return m_n_erased_check_bits; }
  uint32_t n_edges() const { return m_n_edges; }
  uint32_t edge_int_to_out(uint32_t e) const { return m_edge_in_to_out[e]; } // (sic) reference spelling
  uint32_t edge_in_to_out(uint32_t e) const { return m_edge_in_to_out[e]; }
  uint32_t edge_out_to_in(uint32_t e) const { return m_edge_out_to_in[e]; }
  uint32_t out_bit_to_edge(uint32_t b) const { return m_out_bit_to_edge[b]; }
  uint32_t out_edge_to_bit(uint32_t e) const;
  uint32_t in_bit_to_edge(uint32_t b) const { return m_in_bit_to_edge[b]; }
  uint32_t in_edge_to_bit(uint32_t e) const;
  int32_t max_degree_in() const { return m_max_degree_in; }
  int32_t max_degree_out() const { return m_max_degree_out; }

  // whole tables (what the C ABI hands to the device layer)
  const std::vector<uint32_t> &row_ptr() const { return m_out_bit_to_edge; }
  const std::vector<uint32_t> &col_ptr() const { return m_in_bit_to_edge; }
  const std::vector<uint32_t> &in_to_out() const { return m_edge_in_to_out; }
  const std::vector<uint32_t> &out_to_in() const { return m_edge_out_to_in; }
  const std::vector<uint32_t> &out_edge_to_var() const { return m_out_edge_to_in_bit; }

  // true iff offsets are strictly increasing and in range (the decoder constructor's
  // "Incorrect code structure" test, ldpc_decoder_gpu.cu:40-58)
  bool structure_ok() const;

private:
  void parse(const char *buf, size_t len);
  void build_csc();

  std::vector<uint32_t> m_in_bit_to_edge, m_out_bit_to_edge;
  std::vector<uint32_t> m_edge_in_to_out, m_edge_out_to_in;
  std::vector<uint32_t> m_out_edge_to_in_bit;
  int64_t m_n_inputs = 0, m_n_outputs = 0;
  uint32_t m_n_edges = 0;
  int64_t m_n_erased_variables = 0, m_n_erased_check_bits = 0;
  int32_t m_max_degree_in = 0, m_max_degree_out = 0;
};

inline int64_t n_effective_inputs(const ldpc_code &c) { return c.n_inputs() - c.n_erased_inputs(); }
inline int64_t n_effective_outputs(const ldpc_code &c) { return c.n_outputs() - c.n_erased_outputs(); }
// (i-o)/(i-e): reference ldpc_code.cpp:241-254
inline float rate(const ldpc_code &c) {
  return static_cast<float>(c.n_inputs() - c.n_outputs()) / static_cast<float>(c.n_inputs() - c.n_erased_inputs());
}
// H*x on frame-major packed words, LSB-first (scalar equivalent of ldpc_code.cpp:256-286)
void compute_syndrome(const ldpc_code &c, const uint32_t *bits_packed, uint32_t *syndrome_packed);

} // namespace ldpc_b200
