/*
 * ldpc_cpu_fast.c -- TEST INFRASTRUCTURE (CPU baseline timing only).
 *
 * SIMD-over-frames restatement of the reference's OpenCL CN/VN kernels (src/opencl/flood_vec2.cl:174-347):
 * the OpenCL source is float2-vectorised over frame pairs and a CPU OpenCL runtime (PoCL) vectorises work-items
 * further; here the innermost loop runs over the P frame slots (contiguous in memory) and is compiled with
 * -O3 -ffast-math so GCC emits AVX2 code with libmvec expf/logf.  Same message layout, same schedule, same
 * clamp/Taylor switch (flood_vec2.cl:71-116); arithmetic differs from the exact oracle by libmvec accuracy.
 */
#include "ldpc_oracle_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FAST_MAX_P 4096

static inline float phi_abs_fast(float x, float pre) {
  const float xm = x > pre ? x : pre;
  const float e = expf(-xm);
  /* 1 - e^-x without cancellation below 1/8 (expm1 in the reference) */
  float q = xm * (-1.0f / 720.0f) + 1.0f / 120.0f;
  q = xm * q - 1.0f / 24.0f;
  q = xm * q + 1.0f / 6.0f;
  q = xm * q - 0.5f;
  q = xm * q + 1.0f;
  const float den = xm < 0.125f ? xm * q : 1.0f - e;
  const float r = logf((2.0f - den) / den);
  return xm > 5.0f ? 2.0f * e : r;
}

/* flood_backward, flood_vec2.cl:174-232 */
void oracle_fast_backward(oracle_decoder *d) {
  const oracle_code *c = d->c;
  const int64_t P = d->P, M = c->n_checks;
  const float pre = d->pre_threshold;
#pragma omp parallel num_threads(d->n_threads)
  {
    float *S = (float *)aligned_alloc(64, sizeof(float) * FAST_MAX_P);
    uint32_t *X = (uint32_t *)aligned_alloc(64, sizeof(uint32_t) * FAST_MAX_P);
    float *R = (float *)aligned_alloc(64, sizeof(float) * FAST_MAX_P);
#pragma omp for schedule(dynamic, 256)
    for (int64_t o = 0; o < M; o++) {
      const uint32_t a = c->out_bit_to_edge[o], b = c->out_bit_to_edge[o + 1];
      const uint32_t *synw = d->syn + P * (o >> 5);
      const int sh = (int)(o & 31);
      const uint32_t odd = ((b - a) & 1u) << 31;
      for (int64_t f = 0; f < P; f++) {
        S[f] = 0.f;
        X[f] = (((synw[f] >> sh) & 1u) << 31) ^ odd; /* parity seed: syndrome ^ (deg & 1) */
      }
      for (uint32_t e = a; e < b; e++) {
        const float *m = d->msg + P * (int64_t)e;
        const uint32_t *mu = (const uint32_t *)m;
#pragma omp simd
        for (int64_t f = 0; f < P; f++) {
          S[f] += fabsf(m[f]);
          X[f] ^= mu[f];
        }
      }
      for (uint32_t e = a; e < b; e++) {
        float *m = d->msg + P * (int64_t)e;
        uint32_t *mu = (uint32_t *)m;
        const uint32_t *Ru = (const uint32_t *)R;
#pragma omp simd
        for (int64_t f = 0; f < P; f++) R[f] = phi_abs_fast(S[f] - fabsf(m[f]), pre);
#pragma omp simd
        for (int64_t f = 0; f < P; f++) mu[f] = (Ru[f] & 0x7FFFFFFFu) | ((mu[f] ^ X[f]) & 0x80000000u);
      }
    }
    free(S);
    free(X);
    free(R);
  }
}

/* flood_forward / flood_forward_w_final_bits, flood_vec2.cl:236-347 */
void oracle_fast_forward(oracle_decoder *d, int final_bits) {
  const oracle_code *c = d->c;
  const int64_t P = d->P, N = c->n_vars;
  const float pre = d->pre_threshold;
#pragma omp parallel num_threads(d->n_threads)
  {
    float *T = (float *)aligned_alloc(64, sizeof(float) * FAST_MAX_P);
    float *R = (float *)aligned_alloc(64, sizeof(float) * FAST_MAX_P);
    float *D = (float *)aligned_alloc(64, sizeof(float) * FAST_MAX_P);
#pragma omp for schedule(dynamic, 256)
    for (int64_t v = 0; v < N; v++) {
      const uint32_t a = c->in_bit_to_edge[v], b = c->in_bit_to_edge[v + 1];
      const float *l0 = d->llr0 + P * v;
      for (int64_t f = 0; f < P; f++) T[f] = l0[f];
      for (uint32_t ie = a; ie < b; ie++) {
        const float *m = d->msg + P * (int64_t)c->edge_in_to_out[ie];
#pragma omp simd
        for (int64_t f = 0; f < P; f++) T[f] += m[f];
      }
      if (final_bits) {
        uint8_t *fb = d->final_bits + P * v;
        const uint32_t *Tu = (const uint32_t *)T;
        for (int64_t f = 0; f < P; f++) fb[f] = (uint8_t)((~Tu[f]) >> 31);
      }
      for (uint32_t ie = a; ie < b; ie++) {
        float *m = d->msg + P * (int64_t)c->edge_in_to_out[ie];
        uint32_t *mu = (uint32_t *)m;
        const uint32_t *Ru = (const uint32_t *)R, *Du = (const uint32_t *)D;
#pragma omp simd
        for (int64_t f = 0; f < P; f++) {
          D[f] = T[f] - m[f];
          R[f] = phi_abs_fast(fabsf(D[f]), pre);
        }
#pragma omp simd
        for (int64_t f = 0; f < P; f++) mu[f] = (Ru[f] & 0x7FFFFFFFu) | (Du[f] & 0x80000000u);
      }
    }
    free(T);
    free(R);
    free(D);
  }
}
