// common.h -- shared host types of the B200 decoder (C++14).
// Mirrors the names the reference's callers use (h/common.h:13-73): llr types,
// channelType, the sign-convention helpers and the `error` exception.
#pragma once

#include <cstdint>
#include <exception>
#include <string>

namespace ldpc_b200 {

// llr > 0 <=> bit = 1; bit b is sent as +1 / -1 (reference h/common.h:51-59)
inline bool llr_to_bool(float val) { return val > 0.f; }
inline float bool_to_llr(bool val) { return val ? 1.f : -1.f; }

typedef enum _channelType { awgn, bsc, groupGauss, erasure } channelType; // h/common.h:46-49

class error : public std::exception { // h/common.h:61-73
  std::string m_message;

public:
  error() = delete;
  explicit error(const char *p_message) : m_message(p_message) {}
  explicit error(const std::string &p_message) : m_message(p_message) {}
  const char *what() const noexcept override { return m_message.c_str(); }
};

} // namespace ldpc_b200
