/*
 * ldpc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ldpc_oracle.h).
 *
 * CPU restatement of the reference's flood decoder.  It keeps the reference's
 * array layouts, loop order and summation order so that fp32 results agree with
 * the reference CUDA backend to libm accuracy and fp16 results agree bit for bit
 * when given the device phi table.  Nothing in the product path may call this.
 */
#define _GNU_SOURCE
#include "ldpc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ helpers */
static inline float h2f(uint16_t h) {
  _Float16 x;
  memcpy(&x, &h, 2);
  return (float)x;
}
static inline uint16_t f2h(float f) {
  _Float16 x = (_Float16)f; /* round to nearest even */
  uint16_t h;
  memcpy(&h, &x, 2);
  return h;
}
/* half add/sub/mul: one rounding per operation.  Evaluating in float and rounding
 * once to half is exact for + - * because 24 >= 2*11+2 (innocuous double rounding). */
static inline uint16_t hadd(uint16_t a, uint16_t b) { return f2h(h2f(a) + h2f(b)); }
static inline uint16_t hsub(uint16_t a, uint16_t b) { return f2h(h2f(a) - h2f(b)); }
static inline uint16_t hmul(uint16_t a, uint16_t b) { return f2h(h2f(a) * h2f(b)); }

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------- alist: ldpc_code.cpp:45-152 */
typedef struct {
  const char *p, *end;
} cursor;

static void skip_line(cursor *c) { /* istream::ignore(-1u,'\n') */
  while (c->p < c->end && *c->p != '\n') c->p++;
  if (c->p < c->end) c->p++;
}
static int read_i64(cursor *c, int64_t *out) { /* operator>> : skips whitespace incl. newlines */
  while (c->p < c->end && (*c->p == ' ' || *c->p == '\t' || *c->p == '\n' || *c->p == '\r')) c->p++;
  if (c->p >= c->end) return 0;
  char *e;
  long long v = strtoll(c->p, &e, 10);
  if (e == c->p) return 0;
  c->p = e;
  *out = v;
  return 1;
}

static oracle_code *code_from_buffer(const char *buf, size_t len, int *err) {
  cursor c = {buf, buf + len};
  oracle_code *k = (oracle_code *)calloc(1, sizeof(*k));
  if (err) *err = 0;
  /* header extensions "#e=<n>" / "#ec=<n>" : ldpc_code.cpp:52-76 */
  while (c.p < c.end && *c.p == '#') {
    const char *ls = c.p + 1;
    const char *le = ls;
    while (le < c.end && *le != '\n' && *le != ' ' && *le != '\t' && *le != '\r') le++;
    const char *eq = memchr(ls, '=', (size_t)(le - ls));
    if (eq) {
      size_t nl = (size_t)(eq - ls);
      long long v = strtoll(eq + 1, NULL, 10);
      if (nl == 1 && ls[0] == 'e') k->n_erased_vars = v;
      else if (nl == 2 && ls[0] == 'e' && ls[1] == 'c') k->n_erased_checks = v;
    }
    skip_line(&c);
  }
  int64_t M, N;
  if (!read_i64(&c, &M) || !read_i64(&c, &N) || M <= 0 || N <= 0) goto bad; /* checks first: :78-79 */
  k->n_checks = M;
  k->n_vars = N;
  skip_line(&c); /* rest of line 1 */
  skip_line(&c); /* line 2 (max degrees) ignored: :84-85 */
  k->out_bit_to_edge = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(M + 1));
  k->in_bit_to_edge = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(N + 1));
  int64_t e = 0, v;
  for (int64_t i = 0; i < M; i++) { /* :89-97 */
    k->out_bit_to_edge[i] = (uint32_t)e;
    if (!read_i64(&c, &v) || v < 0) goto bad;
    if (v > k->max_deg_check) k->max_deg_check = (int32_t)v;
    e += v;
  }
  k->out_bit_to_edge[M] = (uint32_t)e;
  k->n_edges = e;
  skip_line(&c);
  int64_t e2 = 0;
  for (int64_t i = 0; i < N; i++) { /* :101-110 */
    k->in_bit_to_edge[i] = (uint32_t)e2;
    if (!read_i64(&c, &v) || v < 0) goto bad;
    if (v > k->max_deg_var) k->max_deg_var = (int32_t)v;
    e2 += v;
  }
  k->in_bit_to_edge[N] = (uint32_t)e2;
  skip_line(&c);
  if (e != e2) goto bad; /* :114-117 */
  const int64_t E = e;
  k->in_edge_to_bit = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)E);
  k->out_edge_to_bit = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)E);
  k->edge_out_to_in = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)E);
  k->edge_in_to_out = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)E);
  k->out_edge_to_in_bit = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)E);
  for (int64_t i = 0; i < N; i++)
    for (uint32_t j = k->in_bit_to_edge[i]; j < k->in_bit_to_edge[i + 1]; j++) k->in_edge_to_bit[j] = (uint32_t)i;
  for (int64_t i = 0; i < M; i++)
    for (uint32_t j = k->out_bit_to_edge[i]; j < k->out_bit_to_edge[i + 1]; j++) k->out_edge_to_bit[j] = (uint32_t)i;
  uint32_t *fill = (uint32_t *)calloc((size_t)N, sizeof(uint32_t));
  for (int64_t i = 0; i < M; i++) { /* adjacency, 1-based variable ids: :139-151 */
    for (uint32_t j = k->out_bit_to_edge[i]; j < k->out_bit_to_edge[i + 1]; j++) {
      int64_t col;
      if (!read_i64(&c, &col) || col < 1 || col > N) {
        free(fill);
        goto bad;
      }
      col--;
      uint32_t in_edge = k->in_bit_to_edge[col] + fill[col];
      if (in_edge >= k->in_bit_to_edge[col + 1]) {
        free(fill);
        goto bad;
      }
      fill[col]++;
      k->edge_out_to_in[j] = in_edge;
      k->edge_in_to_out[in_edge] = j;
      k->out_edge_to_in_bit[j] = (uint32_t)col;
    }
    skip_line(&c); /* zero padding on the row is skipped */
  }
  free(fill);
  return k;
bad:
  if (err) *err = 2;
  oracle_code_free(k);
  return NULL;
}

oracle_code *oracle_code_from_alist_text(const char *text, int *err) {
  return code_from_buffer(text, strlen(text), err);
}

oracle_code *oracle_code_from_alist_file(const char *path, int *err) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    if (err) *err = 1; /* "Alist file could not be opened for reading": ldpc_code.cpp:26-29 */
    return NULL;
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)n + 1);
  size_t got = fread(buf, 1, (size_t)n, f);
  fclose(f);
  buf[got] = 0;
  oracle_code *k = code_from_buffer(buf, got, err);
  free(buf);
  return k;
}

void oracle_code_free(oracle_code *c) {
  if (!c) return;
  free(c->out_bit_to_edge);
  free(c->in_bit_to_edge);
  free(c->edge_out_to_in);
  free(c->edge_in_to_out);
  free(c->in_edge_to_bit);
  free(c->out_edge_to_bit);
  free(c->out_edge_to_in_bit);
  free(c);
}

float oracle_code_rate(const oracle_code *c) { /* ldpc_code.cpp:241-254 */
  return (float)(c->n_vars - c->n_checks) / (float)(c->n_vars - c->n_erased_vars);
}

void oracle_compute_syndrome(const oracle_code *c, const uint32_t *bits, uint32_t *syn) {
  const int64_t S = (c->n_checks + 31) >> 5;
  memset(syn, 0, sizeof(uint32_t) * (size_t)S);
  for (int64_t o = 0; o < c->n_checks; o++) {
    uint32_t s = 0;
    for (uint32_t e = c->out_bit_to_edge[o]; e < c->out_bit_to_edge[o + 1]; e++) {
      uint32_t v = c->out_edge_to_in_bit[e];
      s ^= (bits[v >> 5] >> (v & 31)) & 1u;
    }
    syn[o >> 5] |= s << (o & 31);
  }
}

/* -------------------------------------------------------------- phi: flood.cu */
float oracle_phi_abs(float x, float pre_threshold) { /* flood.cu:31-37 */
  const float c_phi_taylor_limit = 5.f;
  float xm = fmaxf(x, pre_threshold);
  float e = expf(-xm);
  return xm > c_phi_taylor_limit ? 2.f * e : logf(-(e + 1.f) / expm1f(-xm));
}

float oracle_phi(float x, float pre_threshold) { /* flood.cu:40-45: magnitude bits | sign bit of x */
  float pa = oracle_phi_abs(fabsf(x), pre_threshold);
  uint32_t a, b;
  memcpy(&a, &pa, 4);
  memcpy(&b, &x, 4);
  a = (a & 0x7FFFFFFFu) | (b & 0x80000000u);
  memcpy(&pa, &a, 4);
  return pa;
}

uint16_t oracle_phi_abs_h(uint16_t xb) { /* flood.cu:21-29 */
  const uint16_t pre = 0x003f, five = 0x4500, two = 0x4000, half_one = 0x3800;
  float x = h2f(xb);
  /* fmax macro (flood.cu:10): (x)>(y)?(x):(y)  -> NaN falls to y */
  uint16_t xm = (x > h2f(pre)) ? xb : pre;
  float fxm = h2f(xm);
  if (fxm > h2f(five)) {
    uint16_t e = f2h(expf(-fxm)); /* hexp: float exp rounded to half */
    return hmul(two, e);
  }
  uint16_t t = f2h(tanhf(h2f(hmul(xm, half_one)))); /* htanh */
  uint16_t l = f2h(logf(h2f(t)));                   /* hlog  */
  return (uint16_t)(l ^ 0x8000u);                   /* unary minus */
}

void oracle_phi_table_h(uint16_t *table) {
  for (uint32_t i = 0; i < 32768; i++) table[i] = oracle_phi_abs_h((uint16_t)i);
}

/* ------------------------------------------------------------------ decoder */
#include "ldpc_oracle_internal.h"

oracle_decoder *oracle_decoder_create(const oracle_code *c, int precision, uint32_t log2_frames, float pre_threshold,
                                      const uint16_t *phi_table_h, int n_threads) {
  if (c->n_vars & 31) return NULL; /* ldpc_decoder_gpu.cu:30-32 */
  oracle_decoder *d = (oracle_decoder *)calloc(1, sizeof(*d));
  d->c = c;
  d->precision = precision;
  d->log2P = log2_frames;
  d->P = 1u << log2_frames;
  d->S = (c->n_checks + 31) >> 5;
  d->pre_threshold = pre_threshold;
  d->n_threads = n_threads > 0 ? n_threads : 1;
  const size_t P = d->P, E = (size_t)c->n_edges, N = (size_t)c->n_vars;
  if (precision == ORACLE_FP32) {
    d->msg = (float *)calloc(E * P, sizeof(float));
    d->llr0 = (float *)calloc(N * P, sizeof(float));
    d->new_llr = (float *)calloc(N * P, sizeof(float));
  } else {
    d->msg_h = (uint16_t *)calloc(E * P, sizeof(uint16_t));
    d->llr0_h = (uint16_t *)calloc(N * P, sizeof(uint16_t));
    d->new_llr_h = (uint16_t *)calloc(N * P, sizeof(uint16_t));
    d->phi_tab = (uint16_t *)malloc(32768 * sizeof(uint16_t));
    if (phi_table_h) memcpy(d->phi_tab, phi_table_h, 32768 * sizeof(uint16_t));
    else oracle_phi_table_h(d->phi_tab);
  }
  d->syn = (uint32_t *)calloc((size_t)d->S * P, sizeof(uint32_t));
  d->new_syn = (uint32_t *)calloc((size_t)d->S * P, sizeof(uint32_t));
  d->final_bits = (uint8_t *)calloc(N * P, 1);
  return d;
}

void oracle_decoder_set_fast(oracle_decoder *d, int fast) { d->fast = fast; }

void oracle_decoder_free(oracle_decoder *d) {
  if (!d) return;
  free(d->msg);
  free(d->llr0);
  free(d->new_llr);
  free(d->msg_h);
  free(d->llr0_h);
  free(d->new_llr_h);
  free(d->phi_tab);
  free(d->syn);
  free(d->new_syn);
  free(d->final_bits);
  free(d);
}

static inline uint16_t phi_h(const oracle_decoder *d, uint16_t x) { /* flood.cu:40-45 (fp16 macros :4-6) */
  return (uint16_t)((d->phi_tab[x & 0x7FFF] & 0x7FFF) | (x & 0x8000));
}

/* transfer_vectors + llr_bsc/llr_biawgn + flood_refill: ldpc_decoder_gpu.cu:218-273, flood.cu:47-75, 297-329 */
void oracle_refill(oracle_decoder *d, int channel, float channel_param, uint32_t cnt, const void *input,
                   const uint32_t *syndromes) {
  const oracle_code *c = d->c;
  const int64_t N = c->n_vars, n_reg = N - c->n_erased_vars, P = d->P, S = d->S;
  if (d->precision == ORACLE_FP32) {
    const float *in = (const float *)input;
    const float f = channel_param;
#pragma omp parallel for num_threads(d->n_threads) schedule(static)
    for (int64_t i = 0; i < N; i++) {
      for (uint32_t j = 0; j < cnt; j++) {
        float llr;
        if (i >= n_reg) llr = 0.f; /* enqueue_clear of the erased tail: ldpc_decoder_gpu.cu:225 */
        else if (channel == ORACLE_BSC) llr = copysignf(f, in[j + (int64_t)cnt * i]); /* flood.cu:58 */
        else llr = in[j + (int64_t)cnt * i] * f;                                     /* flood.cu:73 */
        d->llr0[j + P * i] = llr;
        const float nv = oracle_phi(llr, d->pre_threshold);
        for (uint32_t ie = c->in_bit_to_edge[i]; ie < c->in_bit_to_edge[i + 1]; ie++)
          d->msg[j + P * (int64_t)c->edge_in_to_out[ie]] = nv;
      }
    }
  } else {
    const uint16_t *in = (const uint16_t *)input;
    const uint16_t fh = f2h(channel_param); /* float -> llr_t kernel argument */
#pragma omp parallel for num_threads(d->n_threads) schedule(static)
    for (int64_t i = 0; i < N; i++) {
      for (uint32_t j = 0; j < cnt; j++) {
        uint16_t llr;
        if (i >= n_reg) llr = 0;
        else if (channel == ORACLE_BSC) llr = (uint16_t)((fh & 0x7FFF) | (in[j + (int64_t)cnt * i] & 0x8000));
        else llr = hmul(in[j + (int64_t)cnt * i], fh);
        d->llr0_h[j + P * i] = llr;
        const uint16_t nv = phi_h(d, llr);
        for (uint32_t ie = c->in_bit_to_edge[i]; ie < c->in_bit_to_edge[i + 1]; ie++)
          d->msg_h[j + P * (int64_t)c->edge_in_to_out[ie]] = nv;
      }
    }
  }
  for (uint32_t j = 0; j < cnt; j++)
    for (int64_t w = 0; w < S; w++) d->syn[j + P * w] = syndromes[(int64_t)j * S + w]; /* flood.cu:324-328 */
}

/* flood_backward: flood.cu:77-115 */
void oracle_backward(oracle_decoder *d) {
  if (d->fast && d->precision == ORACLE_FP32) {
    oracle_fast_backward(d);
    return;
  }
  const oracle_code *c = d->c;
  const int64_t P = d->P, M = c->n_checks;
#pragma omp parallel for num_threads(d->n_threads) schedule(dynamic, 64)
  for (int64_t o = 0; o < M; o++) {
    const uint32_t a = c->out_bit_to_edge[o], b = c->out_bit_to_edge[o + 1];
    const uint32_t *synw = d->syn + P * (o >> 5);
    const int sh = (int)(o & 31);
    if (d->precision == ORACLE_FP32) {
      for (int64_t f = 0; f < P; f++) {
        int sb = (int)((synw[f] >> sh) & 1u);
        float ext = 0.f;
        for (uint32_t e = a; e < b; e++) {
          const float m = d->msg[f + P * (int64_t)e];
          ext += fabsf(m);
          sb ^= (signbit(m) == 0);
        }
        for (uint32_t e = a; e < b; e++) {
          float *pm = &d->msg[f + P * (int64_t)e];
          const float m = *pm;
          const float res = oracle_phi_abs(ext - fabsf(m), d->pre_threshold);
          const int is_neg = (signbit(m) != 0) ^ sb;
          *pm = is_neg ? -res : res;
        }
      }
    } else {
      for (int64_t f = 0; f < P; f++) {
        int sb = (int)((synw[f] >> sh) & 1u);
        uint16_t ext = 0;
        for (uint32_t e = a; e < b; e++) {
          const uint16_t m = d->msg_h[f + P * (int64_t)e];
          ext = hadd(ext, (uint16_t)(m & 0x7FFF));
          sb ^= ((m >> 15) == 0);
        }
        for (uint32_t e = a; e < b; e++) {
          uint16_t *pm = &d->msg_h[f + P * (int64_t)e];
          const uint16_t m = *pm;
          const uint16_t pre = hsub(ext, (uint16_t)(m & 0x7FFF));
          /* phi_abs(x) is evaluated on the signed preimage: fmax(x, pre_threshold) clamps negatives */
          uint16_t res;
          if (pre & 0x8000) res = d->phi_tab[0]; /* negative (or -0): clamps to pre_threshold_h */
          else res = d->phi_tab[pre];
          const int is_neg = (int)(m >> 15) ^ sb;
          *pm = is_neg ? (uint16_t)(res ^ 0x8000) : res;
        }
      }
    }
  }
}

/* flood_forward / flood_forward_w_final_bits: flood.cu:117-189 */
void oracle_forward(oracle_decoder *d, int final_bits) {
  if (d->fast && d->precision == ORACLE_FP32) {
    oracle_fast_forward(d, final_bits);
    return;
  }
  const oracle_code *c = d->c;
  const int64_t P = d->P, N = c->n_vars;
#pragma omp parallel for num_threads(d->n_threads) schedule(dynamic, 64)
  for (int64_t v = 0; v < N; v++) {
    const uint32_t a = c->in_bit_to_edge[v], b = c->in_bit_to_edge[v + 1];
    if (d->precision == ORACLE_FP32) {
      for (int64_t f = 0; f < P; f++) {
        float val = d->llr0[f + P * v];
        for (uint32_t ie = a; ie < b; ie++) val += d->msg[f + P * (int64_t)c->edge_in_to_out[ie]];
        if (final_bits) d->final_bits[f + P * v] = (signbit(val) == 0);
        for (uint32_t ie = a; ie < b; ie++) {
          float *pm = &d->msg[f + P * (int64_t)c->edge_in_to_out[ie]];
          *pm = oracle_phi(val - *pm, d->pre_threshold);
        }
      }
    } else {
      for (int64_t f = 0; f < P; f++) {
        uint16_t val = d->llr0_h[f + P * v];
        for (uint32_t ie = a; ie < b; ie++) val = hadd(val, d->msg_h[f + P * (int64_t)c->edge_in_to_out[ie]]);
        if (final_bits) d->final_bits[f + P * v] = ((val >> 15) == 0);
        for (uint32_t ie = a; ie < b; ie++) {
          uint16_t *pm = &d->msg_h[f + P * (int64_t)c->edge_in_to_out[ie]];
          *pm = phi_h(d, hsub(val, *pm));
        }
      }
    }
  }
}

/* check_parity: flood.cu:191-223 */
void oracle_check_parity(oracle_decoder *d, uint8_t *violated) {
  const oracle_code *c = d->c;
  const int64_t P = d->P, M = c->n_checks;
  memset(violated, 0, (size_t)P); /* enqueue_clear: ldpc_decoder_gpu.cu:367 */
#pragma omp parallel for num_threads(d->n_threads) schedule(static)
  for (int64_t f = 0; f < P; f++) {
    uint8_t par = 0;
    for (int64_t o = 0; o < M; o++) {
      uint8_t s = (uint8_t)((d->syn[f + P * (o >> 5)] >> (o & 31)) & 1u);
      for (uint32_t e = c->out_bit_to_edge[o]; e < c->out_bit_to_edge[o + 1]; e++)
        s ^= d->final_bits[f + P * (int64_t)c->out_edge_to_in_bit[e]];
      par |= s;
    }
    violated[f] = par;
  }
}

/* flood_permute_vecs: flood.cu:225-275 (origin slot state moves to dest; final bits are swapped) */
void oracle_permute(oracle_decoder *d, uint32_t n, const uint32_t *origin, const uint32_t *dest) {
  const oracle_code *c = d->c;
  const int64_t P = d->P, N = c->n_vars, E = c->n_edges, S = d->S;
  for (uint32_t t = 0; t < n; t++) {
    const uint32_t o = origin[t], q = dest[t];
    for (int64_t v = 0; v < N; v++) {
      uint8_t bit = d->final_bits[o + P * v];
      d->final_bits[o + P * v] = d->final_bits[q + P * v];
      d->final_bits[q + P * v] = bit;
      if (d->precision == ORACLE_FP32) d->llr0[q + P * v] = d->llr0[o + P * v];
      else d->llr0_h[q + P * v] = d->llr0_h[o + P * v];
    }
    for (int64_t e = 0; e < E; e++) {
      if (d->precision == ORACLE_FP32) d->msg[q + P * e] = d->msg[o + P * e];
      else d->msg_h[q + P * e] = d->msg_h[o + P * e];
    }
    for (int64_t w = 0; w < S; w++) d->syn[q + P * w] = d->syn[o + P * w];
  }
}

/* deinterlace_output: flood.cu:277-295 (LSB-first, all P slots) */
void oracle_deinterlace(oracle_decoder *d, uint32_t *packed) {
  const int64_t P = d->P, W = d->c->n_vars >> 5;
#pragma omp parallel for num_threads(d->n_threads) schedule(static)
  for (int64_t f = 0; f < P; f++)
    for (int64_t w = 0; w < W; w++) {
      uint32_t v = 0;
      for (int i = 0; i < 32; i++) v |= ((uint32_t)d->final_bits[f + P * ((w << 5) + i)]) << i;
      packed[w + W * f] = v;
    }
}

void oracle_read_messages(const oracle_decoder *d, uint32_t slot, float *out) {
  const int64_t P = d->P, E = d->c->n_edges;
  for (int64_t e = 0; e < E; e++)
    out[e] = d->precision == ORACLE_FP32 ? d->msg[slot + P * e] : h2f(d->msg_h[slot + P * e]);
}

void oracle_read_priors(const oracle_decoder *d, uint32_t slot, float *out) {
  const int64_t P = d->P, N = d->c->n_vars;
  for (int64_t v = 0; v < N; v++)
    out[v] = d->precision == ORACLE_FP32 ? d->llr0[slot + P * v] : h2f(d->llr0_h[slot + P * v]);
}

/* prepare_vectors: ldpc_decoder_gpu.cu:199-216 (host gather into the staging stride) */
static void prepare_vectors(const oracle_decoder *d, const void *input, uint32_t in_stride, uint32_t out_stride,
                            uint32_t first, uint32_t count, void *staged) {
  const int64_t n_reg = d->c->n_vars - d->c->n_erased_vars;
  const size_t es = d->precision == ORACLE_FP32 ? 4 : 2;
  for (int64_t i = 0; i < n_reg; i++)
    memcpy((char *)staged + es * (size_t)(i * out_stride), (const char *)input + es * ((size_t)i * in_stride + first),
           es * count);
}

/* decode(): ldpc_decoder_gpu.cu:283-634 */
void oracle_decode(oracle_decoder *d, int channel, float channel_param, uint32_t num_iter_max,
                   uint32_t num_iter_check_parity, uint32_t n_frames, const void *input, const uint32_t *syndromes,
                   uint32_t *results, oracle_report *report, uint32_t *iter_counts, int simd_pairs) {
  if (n_frames == 0) return; /* :293-294 */
  const double t0 = now_s();
  const oracle_code *c = d->c;
  const uint32_t P = d->P;
  const int64_t W = c->n_vars >> 5, S = d->S;
  const uint32_t batch = n_frames < P ? n_frames : P;
  uint32_t next_to_load = batch;
  uint32_t *in_gpu = (uint32_t *)malloc(sizeof(uint32_t) * n_frames);
  uint32_t *iter_start = (uint32_t *)malloc(sizeof(uint32_t) * n_frames);
  uint32_t *iter_end = (uint32_t *)malloc(sizeof(uint32_t) * n_frames);
  uint8_t *violated = (uint8_t *)calloc(P, 1);
  uint8_t *stop = (uint8_t *)calloc(P, 1);
  uint32_t *origin = (uint32_t *)malloc(sizeof(uint32_t) * P);
  uint32_t *dest = (uint32_t *)malloc(sizeof(uint32_t) * P);
  uint32_t *packed = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)W * P);
  const size_t es = d->precision == ORACLE_FP32 ? 4 : 2;
  void *staged = malloc(es * (size_t)c->n_vars * P);
  for (uint32_t i = 0; i < n_frames; i++) iter_start[i] = iter_end[i] = (uint32_t)-1;
  for (uint32_t i = 0; i < batch; i++) in_gpu[i] = i;

  prepare_vectors(d, input, n_frames, batch, 0, batch, staged);
  oracle_refill(d, channel, channel_param, batch, staged, syndromes);

  uint32_t g = 0;
  const double t_iter0 = now_s();
  double t_iter1 = t_iter0;
  for (;;) {
    oracle_backward(d);
    const int check = (g > 0) && (g % num_iter_check_parity == 0); /* :350 */
    if (!check) {
      oracle_forward(d, 0);
    } else {
      oracle_forward(d, 1);
      oracle_check_parity(d, violated);
      uint32_t n_stop = 0;
      memset(stop, 0, P);
      for (uint32_t j = 0; j < batch; j++) { /* :395-403 */
        const uint32_t n_it = g - iter_start[in_gpu[j]];
        if (!violated[j] || n_it >= num_iter_max) {
          n_stop++;
          stop[j] = 1;
          if (iter_end[in_gpu[j]] == (uint32_t)-1) iter_end[in_gpu[j]] = g;
        }
      }
      if (next_to_load == n_frames && n_stop == batch) { /* :414-462 */
        t_iter1 = now_s();
        oracle_deinterlace(d, packed);
        for (uint32_t j = 0; j < batch; j++)
          memcpy(results + (size_t)in_gpu[j] * W, packed + (size_t)j * W, 4 * (size_t)W);
        break;
      }
      uint32_t n_new = n_frames - next_to_load < n_stop ? n_frames - next_to_load : n_stop; /* :464 */
      if (simd_pairs) n_new &= ~1u; /* ldpc_decoder_gpu.cpp:623 */
      if (n_new > 0) {
        uint32_t ctr = 0;
        for (uint32_t i = 0; i < n_new; i++) ctr += stop[i];
        const uint32_t n_swaps = n_new - ctr; /* :487-516 */
        uint32_t o = 0, q = n_new;
        for (uint32_t i = 0; i < n_swaps; i++) {
          while (stop[o]) o++;
          while (!stop[q]) q++;
          origin[i] = o++;
          dest[i] = q++;
        }
        for (uint32_t i = 0; i < n_swaps; i++) {
          uint32_t t = in_gpu[origin[i]];
          in_gpu[origin[i]] = in_gpu[dest[i]];
          in_gpu[dest[i]] = t;
        }
        if (n_swaps) oracle_permute(d, n_swaps, origin, dest); /* :535-549 */
        oracle_deinterlace(d, packed);                          /* :557 */
        for (uint32_t j = 0; j < n_new; j++)
          memcpy(results + (size_t)in_gpu[j] * W, packed + (size_t)j * W, 4 * (size_t)W);
        prepare_vectors(d, input, n_frames, n_new, next_to_load, n_new, staged); /* :588 */
        oracle_refill(d, channel, channel_param, n_new, staged, syndromes + (size_t)next_to_load * S);
        for (uint32_t j = 0; j < n_new; j++) { /* :604-609 */
          in_gpu[j] = next_to_load + j;
          iter_start[next_to_load + j] = g;
        }
        next_to_load += n_new;
      }
    }
    g++;
  }
  if (report) { /* :616-628 */
    report->max_iter = 0;
    report->min_iter = (uint32_t)-1;
    double avg = 0;
    for (uint32_t j = 0; j < n_frames; j++) {
      const uint32_t n = iter_end[j] - iter_start[j];
      if (n > report->max_iter) report->max_iter = n;
      if (n < report->min_iter) report->min_iter = n;
      avg += n;
    }
    report->avg_iter = (float)(avg / n_frames);
    report->iter_time_per_vector = (float)((t_iter1 - t_iter0) / ((double)g * batch));
    report->global_iters = g;
    report->elapsed_s = now_s() - t0;
  }
  if (iter_counts)
    for (uint32_t j = 0; j < n_frames; j++) iter_counts[j] = iter_end[j] - iter_start[j];
  free(in_gpu);
  free(iter_start);
  free(iter_end);
  free(violated);
  free(stop);
  free(origin);
  free(dest);
  free(packed);
  free(staged);
}
