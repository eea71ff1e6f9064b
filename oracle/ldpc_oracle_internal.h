/* ldpc_oracle_internal.h -- TEST INFRASTRUCTURE: decoder state shared by ldpc_oracle.c and ldpc_cpu_fast.c */
#ifndef LDPC_ORACLE_INTERNAL_H
#define LDPC_ORACLE_INTERNAL_H
#include "ldpc_oracle.h"

struct oracle_decoder {
  const oracle_code *c;
  int precision;
  uint32_t log2P, P;
  int64_t S; /* ceil(M/32) */
  float pre_threshold;
  int n_threads;
  /* fp32 state */
  float *msg, *llr0, *new_llr;
  /* fp16 state (raw half bits) */
  uint16_t *msg_h, *llr0_h, *new_llr_h;
  uint16_t *phi_tab; /* [32768] */
  uint32_t *syn, *new_syn;
  uint8_t *final_bits;
  int fast; /* fp32 only: SIMD-over-frames kernels of ldpc_cpu_fast.c (CPU baseline timing) */
};

void oracle_fast_backward(oracle_decoder *d);
void oracle_fast_forward(oracle_decoder *d, int final_bits);
#endif
