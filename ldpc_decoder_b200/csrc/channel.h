// channel.h -- the part of the reference's channel model the decoder reads (C++14).
// Mirrors h/channel.h:18-78 / src/channel.cpp:6-45: the decoder only needs channel(), ref_llr(), factor(), llr().
// Noise injection and capacity integrals are test-data generation (out of scope).
#pragma once

#include "common.h"

#include <cmath>

namespace ldpc_b200 {

class noisy_channel {
public:
  noisy_channel() = default;
  virtual ~noisy_channel() = default;
  virtual float llr(const float) const = 0;
  virtual channelType channel() const = 0;
};

class bsc_channel : public noisy_channel {
  float m_p;
  float m_llr_ref;

public:
  explicit bsc_channel(float p) : m_p(p), m_llr_ref(std::log(1 - p) - std::log(p)) {} // channel.cpp:6-9
  float llr(float val) const override { return val > 0 ? m_llr_ref : -m_llr_ref; }
  float ref_llr() const { return m_llr_ref; }
  float error_probability() const { return m_p; }
  channelType channel() const override { return bsc; }
};

class biawgn_channel : public noisy_channel {
  float m_s;   // noise standard deviation for a +-1 modulation
  float m_snr; // 1/sigma^2

public:
  explicit biawgn_channel(float s) : m_s(s), m_snr(1 / (s * s)) {} // channel.cpp:38-41
  float llr(const float val) const override { return 2 * m_snr * val; }
  float factor() const { return 2 * m_snr; } // h/channel.h:70-73
  float sigma() const { return m_s; }
  channelType channel() const override { return awgn; }
};

} // namespace ldpc_b200
