// ldpc_code.cpp -- alist parser and CSR/CSC construction (see ldpc_code.h).
#include "ldpc_code.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace ldpc_b200 {

namespace {

// Minimal token scanner over the whole file image: the reference reads with iostream
// (ldpc_code.cpp:45-152), which takes tens of seconds for E ~ 3.7M; this is a single pass.
struct scanner {
  const char *p, *end;
  void skip_line() {
    while (p < end && *p != '\n') ++p;
    if (p < end) ++p;
  }
  bool next_int(int64_t &out) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    if (p >= end) return false;
    bool neg = false;
    if (*p == '-') {
      neg = true;
      ++p;
    }
    if (p >= end || *p < '0' || *p > '9') return false;
    int64_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
    out = neg ? -v : v;
    return true;
  }
};

const char *const k_malformed = "PrecomputedCode::init_from_alist_file(): malformed alist file";

} // namespace

ldpc_code::ldpc_code(const std::string &p_alist, bool p_is_filename) {
  if (p_is_filename) {
    std::ifstream f(p_alist.c_str(), std::ios::binary);
    if (!f.good()) throw error("Alist file could not be opened for reading");
    std::string content((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    parse(content.data(), content.size());
  } else {
    parse(p_alist.data(), p_alist.size());
  }
}

ldpc_code::ldpc_code(uint32_t n_checks, uint32_t n_vars, const uint32_t *row_ptr, const uint32_t *col_idx,
                     uint32_t n_erased_vars)
    : m_n_inputs(n_vars), m_n_outputs(n_checks), m_n_erased_variables(n_erased_vars) {
  m_out_bit_to_edge.assign(row_ptr, row_ptr + n_checks + 1);
  m_n_edges = m_out_bit_to_edge[n_checks];
  m_out_edge_to_in_bit.assign(col_idx, col_idx + m_n_edges);
  for (uint32_t e = 0; e < m_n_edges; ++e)
    if (m_out_edge_to_in_bit[e] >= n_vars) throw error(k_malformed);
  for (uint32_t c = 0; c < n_checks; ++c) {
    if (m_out_bit_to_edge[c + 1] < m_out_bit_to_edge[c]) throw error(k_malformed);
    m_max_degree_out = std::max<int32_t>(m_max_degree_out, int32_t(m_out_bit_to_edge[c + 1] - m_out_bit_to_edge[c]));
  }
  build_csc();
}

void ldpc_code::parse(const char *buf, size_t len) {
  scanner s{buf, buf + len};
  // "#name=value" header lines; e = punctured variables, ec = punctured checks (ldpc_code.cpp:52-76)
  while (s.p < s.end && *s.p == '#') {
    const char *tok = s.p + 1;
    const char *tok_end = tok;
    while (tok_end < s.end && !strchr(" \t\r\n", *tok_end)) ++tok_end;
    const char *eq = static_cast<const char *>(memchr(tok, '=', size_t(tok_end - tok)));
    if (eq) {
      std::string name(tok, eq);
      long long v = strtoll(std::string(eq + 1, tok_end).c_str(), nullptr, 10);
      if (name == "e") m_n_erased_variables = v;
      else if (name == "ec") m_n_erased_check_bits = v;
    }
    s.skip_line();
  }
  int64_t M = 0, N = 0;
  if (!s.next_int(M) || !s.next_int(N) || M <= 0 || N <= 0) throw error(k_malformed); // checks first (:78-79)
  m_n_outputs = M;
  m_n_inputs = N;
  s.skip_line(); // remainder of line 1
  s.skip_line(); // line 2 (max degrees) is not used (:84-85)
  m_out_bit_to_edge.resize(size_t(M) + 1);
  int64_t e = 0, v = 0;
  for (int64_t i = 0; i < M; ++i) {
    m_out_bit_to_edge[size_t(i)] = uint32_t(e);
    if (!s.next_int(v) || v < 0) throw error(k_malformed);
    m_max_degree_out = std::max<int32_t>(m_max_degree_out, int32_t(v));
    e += v;
  }
  m_out_bit_to_edge[size_t(M)] = uint32_t(e);
  s.skip_line();
  std::vector<uint32_t> col_deg(size_t(N), 0u);
  int64_t e_check = 0;
  for (int64_t i = 0; i < N; ++i) {
    if (!s.next_int(v) || v < 0) throw error(k_malformed);
    col_deg[size_t(i)] = uint32_t(v);
    e_check += v;
  }
  s.skip_line();
  if (e != e_check || e > int64_t(UINT32_MAX)) throw error(k_malformed); // :114-117
  m_n_edges = uint32_t(e);
  m_out_edge_to_in_bit.resize(m_n_edges);
  for (int64_t i = 0; i < M; ++i) {
    for (uint32_t j = m_out_bit_to_edge[size_t(i)]; j < m_out_bit_to_edge[size_t(i) + 1]; ++j) {
      if (!s.next_int(v) || v < 1 || v > N) throw error(k_malformed); // 1-based variable ids
      m_out_edge_to_in_bit[j] = uint32_t(v - 1);
    }
    s.skip_line(); // zero padding at the end of the row is skipped (:150)
  }
  build_csc();
  // the degree line must agree with the adjacency actually read
  for (int64_t i = 0; i < N; ++i)
    if (m_in_bit_to_edge[size_t(i) + 1] - m_in_bit_to_edge[size_t(i)] != col_deg[size_t(i)]) throw error(k_malformed);
}

// Counting sort of the check-major edges by variable: stable, so a variable's in-edges
// come out in increasing check order (= the reference's tmp_column_indexes fill, :139-151).
void ldpc_code::build_csc() {
  const size_t N = size_t(m_n_inputs);
  m_in_bit_to_edge.assign(N + 1, 0u);
  for (uint32_t e = 0; e < m_n_edges; ++e) ++m_in_bit_to_edge[m_out_edge_to_in_bit[e] + 1];
  for (size_t i = 0; i < N; ++i) {
    m_max_degree_in = std::max<int32_t>(m_max_degree_in, int32_t(m_in_bit_to_edge[i + 1]));
    m_in_bit_to_edge[i + 1] += m_in_bit_to_edge[i];
  }
  std::vector<uint32_t> cursor(m_in_bit_to_edge.begin(), m_in_bit_to_edge.end() - 1);
  m_edge_in_to_out.resize(m_n_edges);
  m_edge_out_to_in.resize(m_n_edges);
  for (uint32_t e = 0; e < m_n_edges; ++e) {
    const uint32_t in_edge = cursor[m_out_edge_to_in_bit[e]]++;
    m_edge_out_to_in[e] = in_edge;
    m_edge_in_to_out[in_edge] = e;
  }
}

uint32_t ldpc_code::out_edge_to_bit(uint32_t e) const {
  return uint32_t(std::upper_bound(m_out_bit_to_edge.begin(), m_out_bit_to_edge.end(), e) - m_out_bit_to_edge.begin() - 1);
}

uint32_t ldpc_code::in_edge_to_bit(uint32_t e) const { return m_out_edge_to_in_bit[m_edge_in_to_out[e]]; }

bool ldpc_code::structure_ok() const {
  // every variable and every check owns at least one edge and offsets stay in range
  for (int64_t i = 0; i < m_n_inputs; ++i) {
    const uint32_t a = m_in_bit_to_edge[size_t(i)];
    if (a >= m_n_edges || (i > 0 && a <= m_in_bit_to_edge[size_t(i) - 1])) return false;
  }
  for (int64_t i = 0; i < m_n_outputs; ++i) {
    const uint32_t a = m_out_bit_to_edge[size_t(i)];
    if (a >= m_n_edges || (i > 0 && a <= m_out_bit_to_edge[size_t(i) - 1])) return false;
  }
  return true;
}

void compute_syndrome(const ldpc_code &c, const uint32_t *bits, uint32_t *syn) {
  const int64_t M = c.n_outputs();
  std::fill(syn, syn + ((M + 31) >> 5), 0u);
  const std::vector<uint32_t> &rp = c.row_ptr(), &var = c.out_edge_to_var();
  for (int64_t o = 0; o < M; ++o) {
    uint32_t s = 0;
    for (uint32_t e = rp[size_t(o)]; e < rp[size_t(o) + 1]; ++e) s ^= (bits[var[e] >> 5] >> (var[e] & 31)) & 1u;
    syn[o >> 5] |= s << (o & 31);
  }
}

} // namespace ldpc_b200
